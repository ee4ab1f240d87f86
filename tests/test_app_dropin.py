"""App-level drop-in ("PccAppEncoder / PccAppDecoder link unchanged"). oracle/Makefile builds the reference's own
applications twice from the sources under /root/reference: unmodified (oracle/_ref/bin/PccAppEncoder, PccAppDecoder) and with
PCCEncoder::encode / PCCDecoder::decode patched (oracle/appkit/patch_reference.py, a copy - the reference tree is untouched) so that
the hot-path calls go through integration/pccb200_shim.cpp into libpccb200.so (…_b200). Both encoders run the same command line -
the CTC parameter values, the pass-through codec stub in place of the external HM process (oracle/appkit/codec_stub.py),
--keepIntermediateFiles - and must produce IDENTICAL files: the V3C bitstream, its checksum file, every YUV frame handed to / returned
by the codec, the 3DMC side files, the conformance logs (atlas / tile / picture / pcframe MD5s: the reference's own pins) and
the reconstructed clouds.

The binaries exist only where /root/reference was present at build time (the build container); they travel to the GPU box with the
snapshot. Without them the tests skip."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
STUB = os.path.join(ROOT, "oracle", "appkit", "codec_stub.py")

# cfg/common/ctc-common.cfg + cfg/sequence/longdress_vox10.cfg (values only; the partitioning / ROI keys are inactive with
# enablePointCloudPartitioning 0) - SURVEY.md 8a-0
CTC_COMMON = {
    "colorTransform": 0, "nnNormalEstimation": 16, "maxNNCountRefineSegmentation": 1024, "voxelDimensionRefineSegmentation": 4,
    "searchRadiusRefineSegmentation": 192, "occupancyResolution": 16, "minPointCountPerCCPatchSegmentation": 16, "maxNNCountPatchSegmentation": 16,
    "surfaceThickness": 4, "maxAllowedDist2RawPointsDetection": 9, "maxAllowedDist2RawPointsSelection": 1, "lambdaRefineSegmentation": 3,
    "minimumImageWidth": 1280, "minimumImageHeight": 1280, "bestColorSearchRange": 0, "numNeighborsColorTransferFwd": 8, "numNeighborsColorTransferBwd": 1,
    "useDistWeightedAverageFwd": 1, "useDistWeightedAverageBwd": 1, "skipAvgIfIdenticalSourcePointPresentFwd": 1, "skipAvgIfIdenticalSourcePointPresentBwd": 1,
    "distOffsetFwd": 4, "distOffsetBwd": 4, "maxGeometryDist2Fwd": 1000, "maxGeometryDist2Bwd": 1000, "maxColorDist2Fwd": 1000, "maxColorDist2Bwd": 1000,
    "maxCandidateCount": 4, "flagGeometrySmoothing": 1, "gridSmoothing": 1, "gridSize": 8, "thresholdSmoothing": 64, "thresholdColorPreSmoothing": 10.0,
    "thresholdColorPreSmoothingLocalEntropy": 4.5, "radius2ColorPreSmoothing": 64, "neighborCountColorPreSmoothing": 64, "flagColorPreSmoothing": 1,
    "enablePointCloudPartitioning": 0, "enhancedOccupancyMapCode": 0, "profileReconstructionIdc": 1,
    "geometry3dCoordinatesBitdepth": 10, "geometryNominal2dBitdepth": 8, "groupOfFramesSize": 32, "minNormSumOfInvDist4MPSelection": 0.33,
    "partialAdditionalProjectionPlane": 0.17, "maxPatchSize": 1024, "numTilesHor": 2, "tileHeightToWidthRatio": 1,
}
CONDITIONS = {   # cfg/condition/ctc-all-intra.cfg + cfg/rate/ctc-r3.cfg ; cfg/condition/ctc-random-access.cfg + cfg/rate/ctc-r5.cfg
    "ai_r3": {"constrainedPack": 0, "globalPatchAllocation": 0, "geometryQP": 24, "attributeQP": 32, "occupancyPrecision": 4},
    "ra_r5": {"globalPatchAllocation": 1, "geometryQP": 16, "attributeQP": 22, "occupancyPrecision": 2},
}


def have_apps():
    return all(os.path.exists(os.path.join(BIN, n)) for n in ("PccAppEncoder", "PccAppEncoder_b200", "PccAppDecoder", "PccAppDecoder_b200"))


def write_ply(path, xyz, rgb):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(xyz))
        f.write("".join("%d %d %d %d %d %d\n" % (*p, *c) for p, c in zip(xyz.tolist(), rgb.tolist())))


def encoder_command(binary, work, out, condition, frames, iterations):
    cfg = os.path.join(work, condition + ".cfg")
    dummy = os.path.join(work, "codec.cfg")       # the codec configuration handed to the stub (it ignores it)
    open(dummy, "w").close()
    with open(cfg, "w") as f:
        for k, v in {**CTC_COMMON, **CONDITIONS[condition]}.items():
            f.write("%s: %s\n" % (k, v))
        for k in ("geometryConfig", "attributeConfig", "occupancyMapConfig", "geometryMPConfig"):
            f.write("%s: %s\n" % (k, dummy))
    os.makedirs(out, exist_ok=True)
    # A non-empty colour-space configuration with an empty colorSpaceConversionPath selects the reference's INTERNAL converter
    # (PCCVideoEncoder.cpp:318-334: "RGB444ToYUV420_8_4" / "YUV420ToYUV444_8_0"); the files only have to exist
    # (PCCEncoderParameters.cpp:832-846). With empty names the attribute frames would reach the codec without any conversion.
    return [os.path.join(BIN, binary), "--config=" + cfg, "--uncompressedDataPath=" + os.path.join(work, "frame_%04d.ply"), "--startFrameNumber=0",
            "--frameCount=%d" % frames, "--nbThread=1", "--colorSpaceConversionConfig=" + dummy, "--inverseColorSpaceConversionConfig=" + dummy,
            "--iterationCountRefineSegmentation=%d" % iterations, "--keepIntermediateFiles=1",
            "--compressedStreamPath=" + os.path.join(out, "s.bin"), "--reconstructedDataPath=" + os.path.join(out, "rec_%04d.ply")] + \
           ["--videoEncoder%sPath=%s" % (k, STUB) for k in ("Occupancy", "Geometry", "Attribute")] + \
           ["--videoEncoder%sCodecId=HMAPP" % k for k in ("Occupancy", "Geometry", "Attribute")]


def run(cmd, log, timeout=300):
    with open(log, "w") as f:
        try:
            return subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, cwd=os.path.dirname(log), timeout=timeout).returncode
        except subprocess.TimeoutExpired:
            return -999


def digest_dir(path):
    """name -> sha256 of every product file (the console log and timing-dependent content excluded)"""
    out = {}
    for n in sorted(os.listdir(path)):
        if n.endswith(".log") and not n.endswith("_log.txt"):
            continue
        with open(os.path.join(path, n), "rb") as f:
            out[n] = hashlib.sha256(f.read()).hexdigest()
    return out


def make_inputs(work, frames, scale):
    for f in range(frames):
        xyz, rgb = synth.figure(scale=scale, seed=3, frame=f)
        write_ply(os.path.join(work, "frame_%04d.ply" % f), xyz, rgb)


@pytest.mark.skipif(not have_apps(), reason="oracle/_ref/bin not built (needs /root/reference at build time)")
def test_unmodified_app_runs_with_the_stub_and_the_b200_app_has_no_cpu_fallback(tmp_path):
    """CPU: the reference encoder runs end to end with the codec stub; the B200 build of the same application refuses to run without a
    device (PCCB200_ERR_NO_DEVICE) instead of computing anything on the CPU"""
    import torch
    work = str(tmp_path)
    make_inputs(work, 1, 0.1)
    assert run(encoder_command("PccAppEncoder", work, os.path.join(work, "ref"), "ai_r3", 1, 2), os.path.join(work, "ref.log")) == 0
    names = os.listdir(os.path.join(work, "ref"))
    assert "s.bin" in names and any(n.endswith("_8bit_p420.yuv") and "geometry" in n for n in names)
    if not torch.cuda.is_available():
        rc = run(encoder_command("PccAppEncoder_b200", work, os.path.join(work, "b200"), "ai_r3", 1, 2), os.path.join(work, "b200.log"))
        assert rc != 0
        with open(os.path.join(work, "b200.log")) as f:
            assert "pccb200: stageA failed with -1" in f.read()


@pytest.mark.gpu
@pytest.mark.skipif(not have_apps(), reason="oracle/_ref/bin not built (needs /root/reference at build time)")
@pytest.mark.parametrize("condition,frames,scale,iterations", [("ai_r3", 2, 0.25, 10), ("ra_r5", 4, 0.2, 6)])
def test_b200_encoder_application_writes_the_reference_files(condition, frames, scale, iterations, tmp_path):
    work = str(tmp_path)
    make_inputs(work, frames, scale)
    assert run(encoder_command("PccAppEncoder", work, os.path.join(work, "ref"), condition, frames, iterations), os.path.join(work, "ref.log")) == 0
    rc = run(encoder_command("PccAppEncoder_b200", work, os.path.join(work, "b200"), condition, frames, iterations), os.path.join(work, "b200.log"))
    if rc != 0:
        with open(os.path.join(work, "b200.log")) as f:
            sys.stderr.write(f.read()[-3000:])
    assert rc == 0
    want, got = digest_dir(os.path.join(work, "ref")), digest_dir(os.path.join(work, "b200"))
    assert sorted(want) == sorted(got), "different sets of output files"
    assert "s.bin" in want and any("_log.txt" in n for n in want) and sum(n.endswith(".yuv") for n in want) >= 6
    bad = [n for n in want if want[n] != got[n]]
    assert bad == [], "files differ between the reference application and its B200 build: %s" % bad


def decoder_command(binary, work, stream, out):
    dummy = os.path.join(work, "codec.cfg")
    os.makedirs(out, exist_ok=True)
    return [os.path.join(BIN, binary), "--compressedStreamPath=" + stream, "--inverseColorSpaceConversionConfig=" + dummy, "--colorTransform=0",
            "--nbThread=1", "--startFrameNumber=0", "--reconstructedDataPath=" + os.path.join(out, "dec_%04d.ply"), "--computeChecksum=1"] + \
           ["--videoDecoder%sPath=%s" % (k, STUB) for k in ("Occupancy", "Geometry", "Attribute")]


@pytest.mark.skipif(not have_apps(), reason="oracle/_ref/bin not built (needs /root/reference at build time)")
def test_reference_applications_round_trip_through_the_codec_stub(tmp_path):
    """CPU: pins the test infrastructure itself - with the pass-through stub (incl. its hand-made HEVC SPS, which the reference's
    parser must read the frame sizes from) the UNMODIFIED decoder reproduces the unmodified encoder's reconstruction and passes its
    own checksum comparison"""
    work = str(tmp_path)
    make_inputs(work, 2, 0.1)
    assert run(encoder_command("PccAppEncoder", work, os.path.join(work, "enc"), "ai_r3", 2, 2), os.path.join(work, "enc.log")) == 0
    assert run(decoder_command("PccAppDecoder", work, os.path.join(work, "enc", "s.bin"), os.path.join(work, "dec")), os.path.join(work, "dec.log")) == 0
    enc, dec = digest_dir(os.path.join(work, "enc")), digest_dir(os.path.join(work, "dec"))
    assert [dec["dec_%04d.ply" % f] for f in range(2)] == [enc["rec_%04d.ply" % f] for f in range(2)]
    with open(os.path.join(work, "dec.log")) as f:
        log = f.read()
    assert "hevcParser= 1280 x 1280 8 bits" in log and "hevcParser= 320 x 320 8 bits" in log


@pytest.mark.gpu
@pytest.mark.skipif(not have_apps(), reason="oracle/_ref/bin not built (needs /root/reference at build time)")
def test_b200_decoder_application_reconstructs_what_the_reference_decoder_does(tmp_path):
    """PccAppDecoder with PCCCodec::generatePointCloud on the B200 (PCCDecoder.cpp:349-351 -> pccb200shim::decodeFrame): the stream of
    the unmodified encoder decodes to the same clouds as with the unmodified decoder, and the decoder's own checksum test against the
    encoder's reconstruction (--computeChecksum) passes for both"""
    work = str(tmp_path)
    make_inputs(work, 2, 0.2)
    assert run(encoder_command("PccAppEncoder", work, os.path.join(work, "enc"), "ai_r3", 2, 6), os.path.join(work, "enc.log")) == 0
    stream = os.path.join(work, "enc", "s.bin")
    outs = {}
    for binary in ("PccAppDecoder", "PccAppDecoder_b200"):
        out = os.path.join(work, binary)
        rc = run(decoder_command(binary, work, stream, out), os.path.join(work, binary + ".log"))
        if rc != 0:
            with open(os.path.join(work, binary + ".log")) as f:
                sys.stderr.write(f.read()[-3000:])
        assert rc == 0, binary
        outs[binary] = {n: h for n, h in digest_dir(out).items() if n.endswith(".ply")}
        assert len(outs[binary]) == 2
    assert outs["PccAppDecoder"] == outs["PccAppDecoder_b200"]
    enc = digest_dir(os.path.join(work, "enc"))
    assert [outs["PccAppDecoder"]["dec_%04d.ply" % f] for f in range(2)] == [enc["rec_%04d.ply" % f] for f in range(2)], "decoder output != encoder reconstruction"
