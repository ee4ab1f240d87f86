#!/usr/bin/env python3
"""Stand-in for the external HM encoder / decoder processes the reference spawns with --videoEncoder*CodecId=HMAPP
(PccLibVideoEncoder/source/PCCHMAppVideoEncoder.cpp:47-106, PccLibVideoDecoder/source/PCCHMAppVideoDecoder.cpp:60-87), so that
PccAppEncoder / PccAppDecoder run end to end in a container without HM. Test infrastructure only.

Lossless pass-through: the "encoder" copies --InputFile to --ReconFile and writes the frames into --BitstreamFile as one NAL unit
whose payload is the base64 text of the zlib-compressed YUV (no zero bytes, so the start-code scan of
PCCVideoBitstream::byteStreamToSampleStream finds exactly one unit and the V3C sample-stream round trip returns the same bytes);
the "decoder" turns such a bitstream back into the YUV file. Selected by the arguments it is called with."""
import base64
import sys
import zlib

args = {}
for a in sys.argv[1:]:
    if a.startswith("--") and "=" in a:
        k, v = a[2:].split("=", 1)
        args[k] = v
HEADER = b"\x00\x00\x00\x01\x40\x01"
if "InputFile" in args:                      # encoder role
    data = open(args["InputFile"], "rb").read()
    open(args["ReconFile"], "wb").write(data)
    open(args["BitstreamFile"], "wb").write(HEADER + base64.b64encode(zlib.compress(data, 1)))
elif "BitstreamFile" in args:                # decoder role
    raw = open(args["BitstreamFile"], "rb").read()
    at = raw.find(b"\x40\x01")
    open(args["ReconFile"], "wb").write(zlib.decompress(base64.b64decode(raw[at + 2:])))
else:
    sys.exit("codec_stub: unknown invocation %r" % sys.argv[1:])
