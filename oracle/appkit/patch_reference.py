#!/usr/bin/env python3
"""Recipe of the app-level drop-in: writes PATCHED COPIES of two reference sources into oracle/_ref/patched/
(git-ignored; the reference tree itself is never touched, no reference source enters the repository):

  PCCEncoder.cpp  PCCEncoder::encode (PccLibEncoder/source/PCCEncoder.cpp:71-424): the hot-path calls are replaced by the three
                  stage functions of integration/pccb200_shim.cpp -
                    generateSegments + initializeContext + placeSegments (:103-110), generateOccupancyMap (:133),
                    generateOccupancyMapVideo (:139), generateBlockToPatchFromOccupancyMapVideo (:168), generateGeometryVideo (:172)
                                                                                                    -> pccb200shim::stageA
                    the generatePointCloud loop (:319-334) + generateAttributeVideo (:341)           -> pccb200shim::stageB1
                    the attribute padding loop (:344-424)                                           -> pccb200shim::stageB2
                    smoothPointCloudPostprocess (:650) and transferColors16bitBP (:657-672) of the post-processing loop
                                                                       -> pccb200shim::smoothGeometry / transferColors16
                  everything else (video compression calls incl. colour conversion, colorPointCloud, bitstream) stays as it is.
  PCCDecoder.cpp  PCCDecoder::decode (PccLibDecoder/source/PCCDecoder.cpp:349-351): the generatePointCloud call of the tile loop
                  -> pccb200shim::decodeFrame; the same two post-processing calls (:404, :416-432) as in the encoder

Every edit is an exact-text replacement that must match exactly once inside the function it targets; the script fails loudly when
the reference text differs (another TMC2 version). oracle/Makefile (targets `apps`, `apps_b200`) compiles the unmodified and the
patched sources into oracle/_ref/bin/PccApp{Encoder,Decoder}[_b200]; tests/test_app_dropin.py compares their outputs.
Usage: patch_reference.py <reference root> <output dir>"""
import os
import sys


def function_span(text, signature):
    """[start, end) of the function body that starts at `signature` (brace matching; the reference has no braces in strings there)"""
    start = text.index(signature)
    i = text.index("{", start)
    depth = 0
    while True:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return start, i + 1
        i += 1


def replace_once(body, old, new, what):
    if body.count(old) != 1:
        sys.exit("patch_reference: %s: expected exactly one match of %r, found %d" % (what, old[:60], body.count(old)))
    return body.replace(old, new)


def patch_postprocessing(body, session, what):
    """the post-reconstruction chain of the CTC path, the same three calls in encode() and decode()"""
    body = replace_once(body, "        smoothPointCloudPostprocess( reconstruct, params_.colorTransform_, ppSEIParams, partition );\n", """        {  // grid-based geometry smoothing on the B200
          const int b200rc = pccb200shim::smoothGeometry( %s, reconstruct, partition, ppSEIParams.gridSize_, ppSEIParams.thresholdSmoothing_ );
          if ( b200rc != 0 ) {
            fprintf( stderr, "pccb200: smoothGeometry failed with %%d\\n", b200rc );
            exit( -1 );
          }
        }
""" % session, what + ": smoothPointCloudPostprocess")
    a = body.index("            tempFrameBuffer.transferColors16bitBP( reconstruct,")
    b = body.index(";", body.index("// maxColorDist2Bwd", a)) + 1   # the end of that call statement
    if body.count("            tempFrameBuffer.transferColors16bitBP( reconstruct,") != 1:
        sys.exit("patch_reference: %s: transferColors16bitBP call not unique" % what)
    body = body[:a] + """            if ( params_.attrTransferFilterType_ == 1 && !isAttributes444 ) {  // colour transfer onto the smoothed cloud on the B200
              const int b200rc = pccb200shim::transferColors16( %s, tempFrameBuffer, reconstruct );
              if ( b200rc != 0 ) {
                fprintf( stderr, "pccb200: transferColors16 failed with %%d\\n", b200rc );
                exit( -1 );
              }
            } else
""" % session + body[a:b] + body[b:]
    return body


def patch_encoder(src):
    a, b = function_span(src, "int PCCEncoder::encode( const PCCGroupOfFrames& sources, PCCContext& context, PCCGroupOfFrames& reconstructs ) {")
    body = src[a:b]
    body = replace_once(body, """  generateSegments( sources, context );
""", """  {  // a1-a21 on the B200 (integration/pccb200_shim.cpp): segmentation, packing, occupancy map + video, block-to-patch, geometry video
    const int b200rc = pccb200shim::stageA( g_pccb200, params_, sources, context );
    if ( b200rc != 0 ) {
      fprintf( stderr, "pccb200: stageA failed with %d\\n", b200rc );
      exit( b200rc == PCCB200_ERR_CANVAS ? 180 : -1 );
    }
  }
""", "generateSegments")
    body = replace_once(body, "  params_.initializeContext( context );\n", "  // (params_.initializeContext( context ) is called by stageA, between segmentation and placement as here)\n", "initializeContext")
    body = replace_once(body, "  placeSegments( sources, context );\n", "", "placeSegments")
    body = replace_once(body, "  generateOccupancyMap( context, true );\n", "", "generateOccupancyMap")
    body = replace_once(body, "  generateOccupancyMapVideo( sources, context );\n", "", "generateOccupancyMapVideo")
    body = replace_once(body, "    generateBlockToPatchFromOccupancyMapVideo( context, params_.occupancyResolution_, params_.occupancyPrecision_ );\n",
                        "    // (block-to-patch of the lossless occupancy video: delivered by stageA)\n", "generateBlockToPatchFromOccupancyMapVideo")
    body = replace_once(body, "  generateGeometryVideo( sources, context );\n", "", "generateGeometryVideo")
    body = replace_once(body, "  for ( size_t frameIdx = 0; frameIdx < context.size(); frameIdx++ ) {\n    auto& frame = context[frameIdx];\n",
                        """  {  // a22-a25 on the B200: generatePointCloud of every frame + colour transfer + attribute frames before padding
    const int b200rc = pccb200shim::stageB1( g_pccb200, params_, context, reconstructs, partitions );
    if ( b200rc != 0 ) {
      fprintf( stderr, "pccb200: stageB1 failed with %d\\n", b200rc );
      exit( -1 );
    }
  }
  for ( size_t frameIdx = 0; false && frameIdx < context.size(); frameIdx++ ) {  // (the reference loop, not executed)
    auto& frame = context[frameIdx];
""", "generatePointCloud loop")
    body = replace_once(body, "    generateAttributeVideo( sources, reconstructs, context, params_ );\n", "", "generateAttributeVideo")
    body = replace_once(body, "    if ( params_.attributeBGFill_ < 3 ) {\n", """    if ( params_.attributeBGFill_ == 1 ) {  // a26 on the B200: push-pull padding + group dilation of the attribute frames
      const int b200rc = pccb200shim::stageB2( g_pccb200, context );
      if ( b200rc != 0 ) {
        fprintf( stderr, "pccb200: stageB2 failed with %d\\n", b200rc );
        exit( -1 );
      }
    } else if ( params_.attributeBGFill_ < 3 ) {
""", "attribute padding loop")
    body = patch_postprocessing(body, "g_pccb200", "PCCEncoder::encode")
    head = src[:a]
    marker = '#include "PCCEncoder.h"\n'
    if head.count(marker) != 1:
        sys.exit("patch_reference: include anchor of PCCEncoder.cpp not found")
    head = head.replace(marker, marker + '#include "pccb200_shim.h"\nstatic pccb200shim::Session g_pccb200;  // one B200 context + GOF handle per encoder process\n')
    return head + body + src[b:]


def patch_decoder(src):
    a, b = function_span(src, "int PCCDecoder::decode( PCCContext& context, PCCGroupOfFrames& reconstructs, int32_t atlasIndex = 0 ) {") \
        if "int PCCDecoder::decode( PCCContext& context, PCCGroupOfFrames& reconstructs, int32_t atlasIndex = 0 ) {" in src else \
        function_span(src, "int PCCDecoder::decode( PCCContext& context, PCCGroupOfFrames& reconstructs, int32_t atlasIndex ) {")
    body = src[a:b]
    body = replace_once(body, "      generatePointCloud( tileReconstrct, context, frameIdx, tileIdx, gpcParams, partition, true );\n",
                        """      {  // PCCCodec::generatePointCloud on the B200 (integration/pccb200_shim.cpp)
        const int b200rc = pccb200shim::decodeFrame( g_pccb200dec, context, frameIdx, context.getOccupancyPrecision(), tileReconstrct, partition );
        if ( b200rc != 0 ) {
          fprintf( stderr, "pccb200: decodeFrame failed with %d\\n", b200rc );
          exit( -1 );
        }
      }
""", "generatePointCloud")
    body = patch_postprocessing(body, "g_pccb200dec", "PCCDecoder::decode")
    head = src[:a]
    marker = '#include "PCCDecoder.h"\n'
    if head.count(marker) != 1:
        sys.exit("patch_reference: include anchor of PCCDecoder.cpp not found")
    head = head.replace(marker, marker + '#include "pccb200_shim.h"\nstatic pccb200shim::Session g_pccb200dec;\n')
    return head + body + src[b:]


if __name__ == "__main__":
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for rel, fn in (("source/lib/PccLibEncoder/source/PCCEncoder.cpp", patch_encoder), ("source/lib/PccLibDecoder/source/PCCDecoder.cpp", patch_decoder)):
        with open(os.path.join(ref, rel)) as f:
            text = f.read()
        with open(os.path.join(out, os.path.basename(rel)), "w") as f:
            f.write(fn(text))
    print("patched copies written to", out)
