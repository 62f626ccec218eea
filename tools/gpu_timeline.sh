#!/bin/bash
# tools/timeline_frame.py: one isolated frame, one frame without PDL, and the last frame of 40 / 300 back-to-back frames (clocks under load).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/timeline_frame.py > gpurun_out/timeline_pdl.txt 2>&1; tail -3 gpurun_out/timeline_pdl.txt
python tools/timeline_frame.py --no-pdl > gpurun_out/timeline_nopdl.txt 2>&1; tail -3 gpurun_out/timeline_nopdl.txt
python tools/timeline_frame.py --burst 40 > gpurun_out/timeline_burst40.txt 2>&1; tail -30 gpurun_out/timeline_burst40.txt
python tools/timeline_frame.py --burst 300 --replays 2 > gpurun_out/timeline_burst300.txt 2>&1; tail -3 gpurun_out/timeline_burst300.txt
