"""CPU-only checks of the host side: C-ABI surface, the state_dict contract, error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rerevst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rrv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds():
    from rerevst_code_b200 import _lib
    assert _header_functions() == sorted(_lib.SIGNATURES)


def test_library_loads_and_exports_every_declared_symbol():
    from rerevst_code_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(handle, name), name
    assert _lib.lib().rrv_abi_version() == _lib.ABI_VERSION == 3
    assert _lib.lib().rrv_launch_count() == 0


def test_ctypes_structs_match_header_layout(tmp_path):
    """sizeof/offsetof of the ctypes mirrors == what a C compiler makes of include/rerevst_b200.h."""
    import subprocess
    from rerevst_code_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rerevst_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(rrv_epilogue), sizeof(rrv_conv), '
                   'offsetof(rrv_conv, ep), offsetof(rrv_conv, out_mode), offsetof(rrv_epilogue, norm2));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(_lib.Epilogue), ctypes.sizeof(_lib.Conv), _lib.Conv.ep.offset,
                   _lib.Conv.out_mode.offset, _lib.Epilogue.norm2.offset]


def test_state_dict_contract_matches_reference_keys():
    from rerevst_code_b200.style_network_global import TransformerNet
    from rerevst_code_b200.weights import key_shapes, synthetic_state_dict, check_state_dict
    net = TransformerNet()
    sd = net.state_dict()
    want = key_shapes()
    assert list(sd.keys()).sort() == list(want.keys()).sort() and len(sd) == 107
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    syn = synthetic_state_dict(0)
    net.load_state_dict(syn, strict=True)
    check_state_dict(syn)
    bad = dict(syn)
    bad.pop("Vgg19.slice1.0.weight")
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad, strict=True)
    with pytest.raises(RuntimeError):
        check_state_dict(bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference not present")
def test_state_dict_keys_equal_live_reference(state_dict):
    from oracle.make_golden import import_reference
    g = import_reference("style_network_global", "test")
    ref = g.TransformerNet().state_dict()
    assert set(ref.keys()) == set(state_dict.keys())
    for k in ref:
        assert tuple(ref[k].shape) == tuple(state_dict[k].shape)


def test_no_cpu_fallback():
    from rerevst_code_b200.style_network_global import TransformerNet
    from rerevst_code_b200.framework import Stylization
    from rerevst_code_b200.loss_networks import warp
    net = TransformerNet()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError):
        Stylization({}, cuda=False)
    with pytest.raises(RuntimeError):
        warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))


def test_add_before_clean_fails_like_reference():
    from rerevst_code_b200.style_network_global import TransformerNet
    with pytest.raises(AttributeError):
        TransformerNet().add(torch.zeros(1, 3, 16, 16))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rerevst-code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import / dynamic import / path reference of the oracle package anywhere in the product (prose may name it)
                bad = re.findall(r"^\s*(?:from|import)\s+oracle\b|import_module\(\s*[\"']oracle|__import__\(\s*[\"']oracle|[\"'/]oracle/",
                                 text, flags=re.M)
                assert not bad, (dirpath, f, bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/train"), reason="reference not present")
def test_fake_flow_equals_live_reference():
    """TemporalLoss.GenerateFakeFlow (train/loss_networks.py:71-86): host-side numpy/cv2 synthesis drawing from
    np.random / random in the reference's order -- equal seeds give the reference's flow bit for bit."""
    import random
    import sys

    import numpy as np
    import torch
    from oracle.make_golden import import_reference
    from rerevst_code_b200.loss_networks import TemporalLoss
    ln = import_reference("loss_networks", "train")
    for (h, w) in ((256, 320), (512, 512)):
        np.random.seed(0); random.seed(0)
        ref = ln.TemporalLoss().GenerateFakeFlow(h, w)
        np.random.seed(0); random.seed(0)
        got = TemporalLoss().GenerateFakeFlow(h, w)
        assert got.dtype == torch.float32 and tuple(got.shape) == (2, h, w)
        assert torch.equal(got, ref)
