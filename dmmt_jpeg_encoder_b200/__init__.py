"""dmmt_jpeg_encoder_b200 -- B200-native (sm_100a) encode hot path of dmmt-jpeg-encoder.

`reference_api` mirrors the reference crate's public interface (convert_ppm_to_jpeg, CLIParser,
PPMImageReader, JpegImageWriter, ...); `encoder` exposes the C ABI objects (Context, Plan, Batch);
`sharded` is the one-process-per-GPU MCU-row sharding of a single image.  All compute happens in
libdmmt_cuda.so (csrc/, include/dmmt_cuda.h); there is no CPU encode path in this package.
"""
from . import _ffi
from ._ffi import DmmtError
from .encoder import Batch, Context, Options, PinnedBuffer, Plan
from .reference_api import (Arguments, ChromaSubsamplingPreset, CLIParser, Image, JpegImageWriter,
                            JpegTransformationOptions, PPMImageReader, QuantizationTablePreset,
                            convert_ppm_to_jpeg, main)

__all__ = ["Arguments", "Batch", "CLIParser", "ChromaSubsamplingPreset", "Context", "DmmtError", "Image",
           "JpegImageWriter", "JpegTransformationOptions", "Options", "PPMImageReader", "PinnedBuffer", "Plan",
           "QuantizationTablePreset", "convert_ppm_to_jpeg", "main", "_ffi"]
