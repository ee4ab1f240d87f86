#!/usr/bin/env python3
"""Stand-in for the external HM encoder / decoder processes the reference spawns with --videoEncoder*CodecId=HMAPP
(PccLibVideoEncoder/source/PCCHMAppVideoEncoder.cpp:47-106, PccLibVideoDecoder/source/PCCHMAppVideoDecoder.cpp:60-87), so that
PccAppEncoder / PccAppDecoder run end to end in a container without HM. Test infrastructure only.

Lossless pass-through: the "encoder" copies --InputFile to --ReconFile and writes into --BitstreamFile two NAL units: a real HEVC
sequence parameter set with the picture size (the reference's decoder parses it) and one unit whose payload is the base64 text of
the zlib-compressed YUV (no zero bytes, so the start-code scan of PCCVideoBitstream::byteStreamToSampleStream finds exactly these
units and the V3C sample-stream round trip returns the same payload); the "decoder" turns such a bitstream back into the YUV file.
Selected by the arguments it is called with."""
import base64
import sys
import zlib

args = {}
for a in sys.argv[1:]:
    if a.startswith("--") and "=" in a:
        k, v = a[2:].split("=", 1)
        args[k] = v
class Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        self.b += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def bytes(self):
        bits = self.b + [1]                       # rbsp_stop_one_bit
        bits += [0] * (-len(bits) % 8)
        return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))


def sps_nal(width, height, bit_depth, chroma):
    """A syntactically complete HEVC sequence parameter set (ITU-T H.265 7.3.2.2) carrying the picture size: the decoder side of the
    reference reads width / height / bit depth of a video sub-stream from its SPS (PccLibHevcParser getVideoSize) before it spawns
    the decoder process. Field values are chosen so that no start-code-like byte pattern occurs."""
    b = Bits()
    b.u(4, 0), b.u(3, 0), b.u(1, 1)               # vps id, max_sub_layers_minus1, temporal_id_nesting
    b.u(2, 0), b.u(1, 0), b.u(5, 1)               # profile_tier_level: profile space, tier, profile idc (Main)
    b.u(32, 0x60404041)                           # profile compatibility flags (Main, Main 10; the other set bits only keep every byte non-zero)
    b.u(4, 0b1011)                                # progressive, interlaced, non-packed, frame-only
    b.u(44, 0x00100100101)                        # 43 reserved bits + 1 (no run of 16 zero bits)
    b.u(8, 120)                                   # level 4
    b.ue(0)                                       # sps id
    b.ue({"400": 0, "420": 1, "422": 2, "444": 3}[chroma])
    if chroma == "444":
        b.u(1, 0)
    b.ue(width), b.ue(height), b.u(1, 0)          # size, no conformance window
    b.ue(bit_depth - 8), b.ue(bit_depth - 8)
    b.ue(4), b.u(1, 1), b.ue(4), b.ue(0), b.ue(0)  # poc bits, sub-layer ordering info
    b.ue(0), b.ue(3), b.ue(0), b.ue(3), b.ue(1), b.ue(1)   # coding / transform block sizes, hierarchy depths
    b.u(1, 0), b.u(1, 0), b.u(1, 0), b.u(1, 0)    # scaling list, amp, sao, pcm
    b.ue(0), b.u(1, 0), b.u(1, 0), b.u(1, 0)      # short-term ref pic sets, long-term, temporal mvp, strong intra smoothing
    b.u(1, 0), b.u(1, 0)                          # vui, extension
    body = b.bytes()
    assert b"\x00\x00\x00" not in body and b"\x00\x00\x01" not in body and b"\x00\x00\x02" not in body and b"\x00\x00\x03" not in body
    return b"\x00\x00\x00\x01\x42\x01" + body


PAYLOAD = b"\x00\x00\x00\x01\x40\x01"      # (a VPS-typed unit: the reference's parser skips it)
if "InputFile" in args:                      # encoder role
    data = open(args["InputFile"], "rb").read()
    open(args["ReconFile"], "wb").write(data)
    sps = sps_nal(int(args["SourceWidth"]), int(args["SourceHeight"]), int(args.get("InputBitDepth", 8)), args.get("InputChromaFormat", "420"))
    open(args["BitstreamFile"], "wb").write(sps + PAYLOAD + base64.b64encode(zlib.compress(data, 1)))
elif "BitstreamFile" in args:                # decoder role
    raw = open(args["BitstreamFile"], "rb").read()
    at = raw.rfind(b"\x40\x01")
    open(args["ReconFile"], "wb").write(zlib.decompress(base64.b64decode(raw[at + 2:])))
else:
    sys.exit("codec_stub: unknown invocation %r" % sys.argv[1:])
