"""The drop-in boundary, exercised from the reference's side: integration/pccb200_shim.cpp is the code a maintainer adds to
PccLibEncoder (it uses the reference's own classes and calls only the C ABI). oracle/Makefile compiles it against the reference
headers and links it with the reference objects and libpccb200.so (target `shim`).

CPU: it builds, links and - without a GPU - reports PCCB200_ERR_NO_DEVICE through the reference's harness instead of computing
anything.  GPU: the reference's data structures (PCCPatch, PCCFrameContext, PCCImage, PCCPointSet3), filled through the shim, hold
exactly what the unmodified reference stages leave in them, for all-intra and random-access packing."""
import os
import subprocess

import numpy as np
import pytest

import bindings
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim():
    if os.path.isdir("/root/reference/source/lib"):   # the build container: (re)build from the sources where they lie
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j", "8", "ref", "shim"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not bindings.Shim.available():
        pytest.skip("oracle/_ref/libtmc2shim.so not built (needs /root/reference)")
    return bindings.Shim()


def test_shim_compiles_links_and_refuses_without_gpu(shim):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    frames = [synth.sphere(radius=10, center=40)]
    prm = bindings.ctc_seg_params(bits=10, iterations=2, weight=(1.0, 1.0, 1.0))
    got, code = shim.encode_gof(frames, prm, stop_after=1)
    assert code == -1 and got == []   # PCCB200_ERR_NO_DEVICE: nothing is computed on the CPU


@pytest.mark.gpu
@pytest.mark.parametrize("ra", [0, 1])
def test_gpu_shim_fills_reference_structures_like_the_reference(ra, shim):
    frames = [synth.figure(scale=0.12, seed=4, frame=f) for f in range(3)] + [synth.double_sheet(n_side=32, seed=2)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=(1.0, 1.0, 1.0))   # (both sides compute the axis weights themselves)
    prm.global_patch_allocation = ra
    want, _ = shim.ref.encode_gof(frames, prm, occupancy_precision=2 if ra else 4)
    got, code = shim.encode_gof(frames, prm, occupancy_precision=2 if ra else 4)
    assert code == 0
    assert bindings.compare_gof(got, want) == []


def test_decoder_side_binding_refuses_without_gpu(shim):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    frames = [synth.sphere(radius=10, center=40)]
    prm = bindings.ctc_seg_params(bits=10, iterations=2, weight=(1.0, 1.0, 1.0))
    got, code = shim.decode_gof(frames, prm)
    assert code == -1   # PCCB200_ERR_NO_DEVICE from pccb200shim::decodeFrame: the reference stages ran, nothing was reconstructed on the CPU
