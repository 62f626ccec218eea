"""B200-native ReReVST per-frame stylization path."""
