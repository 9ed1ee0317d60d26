"""Host-side mirror of the reference crate's public encode interface, over the CUDA C ABI.

Same names, argument meaning and error behaviour as dmmt-jpeg-encoder (citations are relative to
/root/reference/src):

  convert_ppm_to_jpeg(arguments)          lib.rs:59-77
  CLIParser / Arguments                   cli.rs:12-180, lib.rs:33-40
  PPMImageReader(reader).read_image()     image/reader/ppm.rs:9-251
  Image                                   image.rs:7-11
  ChromaSubsamplingPreset                 image/subsampling.rs:11-46
  QuantizationTablePreset                 image/writer/jpeg/quantization_tables.rs:232-284
  JpegTransformationOptions               image/writer/jpeg.rs:25-39
  JpegImageWriter(writer, image, options, threadpool).write_image()   image/writer/jpeg.rs:41-75
  Error (+ Display strings)               error.rs:4-98

The body of write_image is ONE call into the CUDA library (dmmt_encode); everything before it
(file handling, the ASCII P3 parse) is host I/O.  There is no CPU encode path here.
"""
from __future__ import annotations

import argparse
import enum
import os
import re
import sys
from dataclasses import dataclass

import numpy as np

from . import _ffi as F
from .encoder import Context, Options


# ------------------------------------------------------------------------------------- errors
class Error(Exception):
    """crate::Error (error.rs:4-23); str() is the reference's Display text (error.rs:25-98)."""


class PPMFileDoesNotContainRequiredToken(Error):
    def __init__(self, token_name):
        super().__init__(f"Expected token '{token_name}' not found in PPM file")


class ParsingOfTokenFailed(Error):
    def __init__(self, token_name):
        super().__init__(f"Parsing of token '{token_name}' failed")


class IncompletePixelParsed(Error):
    def __init__(self, n):
        self.number_of_tokens_parsed = n
        super().__init__(f"Incomplete pixel parsed. Expected 3 components, but got {n}.")


class MismatchOfSizeBetweenHeaderAndValues(Error):
    def __init__(self):
        super().__init__("Nubmer of pixels do not match the size, provided in header")  # sic (error.rs:46)


class UnableToOpenInputFileForReading(Error):
    def __init__(self, path, err):
        super().__init__(f"Unable to open input file '{path}' for reading: {err}")


class UnableToOpenOutputFileForWriting(Error):
    def __init__(self, path, err):
        super().__init__(f"Unable to open output file '{path}' for writing: {err}")


class FailedToWriteImageData(Error):
    def __init__(self):
        super().__init__("Failed to write image data")


class HuffmanSymbolNotPresentInTranslator(Error):
    def __init__(self, symbol="?", translator="?"):
        super().__init__(f"Huffman symbol '{symbol}' not present in {translator} translator")


class ReferencePanic(RuntimeError):
    """Conditions on which the reference panics instead of returning an Error."""


P3_HEADER_TOKEN_NAME = "P3 Header"
WIDTH_HEADER_TOKEN_NAME = "Width Header"
HEIGHT_HEADER_TOKEN_NAME = "Height Header"
MAX_VALUE_HEADER_TOKEN_NAME = "Max Value Header"
COLOR_COMPONENT_VALUE_TOKEN_NAME = "Color Component Value"


# -------------------------------------------------------------------------------------- enums
class ChromaSubsamplingPreset(enum.Enum):
    P444 = F.P444
    P422 = F.P422
    P420 = F.P420

    def horizontal_rate(self) -> int:
        return 1 if self is ChromaSubsamplingPreset.P444 else 2

    def vertical_rate(self) -> int:
        return 2 if self is ChromaSubsamplingPreset.P420 else 1


class QuantizationTablePreset(enum.Enum):
    Specification = 0
    Flat = 1
    MSSIMKodakTuned = 2
    PSNRHVSNKodakTuned = 3
    DCTunePerceptualOptimization = 4
    AVisualDetectionModel = 5
    AnImprovedDetectionModel = 6


# CLI spellings + aliases (quantization_tables.rs:260-283); the numeric aliases skip 3 and 5
QUANTIZATION_TABLE_CLI_NAMES = {
    "Specification": QuantizationTablePreset.Specification, "Spec": QuantizationTablePreset.Specification,
    "Default": QuantizationTablePreset.Specification, "0": QuantizationTablePreset.Specification,
    "Flat": QuantizationTablePreset.Flat, "1": QuantizationTablePreset.Flat,
    "MSSIM-Kodak-Tuned": QuantizationTablePreset.MSSIMKodakTuned, "2": QuantizationTablePreset.MSSIMKodakTuned,
    "PSNR-HVS-N-Kodak-Tuned": QuantizationTablePreset.PSNRHVSNKodakTuned, "4": QuantizationTablePreset.PSNRHVSNKodakTuned,
    "DCTune-Perceptual-Optimization": QuantizationTablePreset.DCTunePerceptualOptimization,
    "6": QuantizationTablePreset.DCTunePerceptualOptimization,
    "A-visual-detection-model": QuantizationTablePreset.AVisualDetectionModel, "7": QuantizationTablePreset.AVisualDetectionModel,
    "An-improved-detection-model": QuantizationTablePreset.AnImprovedDetectionModel,
    "8": QuantizationTablePreset.AnImprovedDetectionModel,
}


# -------------------------------------------------------------------------------------- image
class Image:
    """Image<f32> (image.rs:7-11): `dots` are v/max-normalised f32 RGB, [height, width, 3].

    When built by PPMImageReader the raw samples and max value are kept as well; the device then
    performs the identical IEEE `v as f32 / max as f32` itself (color.rs:45-53) on 1/4 of the bytes.
    """

    def __init__(self, width: int, height: int, dots: np.ndarray | None = None, *,
                 samples: np.ndarray | None = None, max_value: int | None = None):
        self.width, self.height = int(width), int(height)
        self.samples, self.max_value = samples, max_value
        self._dots = None if dots is None else np.ascontiguousarray(dots, dtype=np.float32).reshape(height, width, 3)
        if dots is None and samples is None:
            raise ValueError("Image needs dots or samples")

    @property
    def dots(self) -> np.ndarray:
        if self._dots is None:
            self._dots = self.samples.astype(np.float32) / np.float32(self.max_value)  # color.rs:45-53
        return self._dots


_COMMENT = re.compile(rb"#[^\n]*(?:\n|$)")
_RUST_WS = re.compile(rb"[ \t\n\x0c\r]+")
_U16_TOKEN = re.compile(rb"\+?[0-9]+$")


def _parse_u16(token: bytes, name: str) -> int:
    if not _U16_TOKEN.match(token) or int(token) > 65535:
        raise ParsingOfTokenFailed(name)
    return int(token)


class PPMImageReader:
    """ASCII P3 reader with the reference's token rules (ppm.rs:41-251): `#` starts a comment that
    runs through the next newline ANYWHERE (even inside a token), tokens are split on ASCII
    whitespace, every number must parse as u16."""

    def __init__(self, reader, native: bool = False, threads: int = 1):
        """native=True: tokenise with the library's host-side reader (dmmt_ppm_parse, csrc/ppm_parse.hpp) -- the same
        rules and errors, 50-100x faster than the pure-Python mirror below (which stays the independent check of it)."""
        self.reader, self.native, self.threads = reader, native, threads

    def _read_native(self, data: bytes) -> Image:
        st, detail, width, height, max_value, samples = parse_ppm_native(data, self.threads)
        names = (P3_HEADER_TOKEN_NAME, WIDTH_HEADER_TOKEN_NAME, HEIGHT_HEADER_TOKEN_NAME, MAX_VALUE_HEADER_TOKEN_NAME,
                 COLOR_COMPONENT_VALUE_TOKEN_NAME)
        if st == 1:
            raise PPMFileDoesNotContainRequiredToken(names[detail])
        if st == 2:
            raise ParsingOfTokenFailed(names[detail])
        if st == 3:
            raise IncompletePixelParsed(detail)
        if st == 4:
            raise MismatchOfSizeBetweenHeaderAndValues()
        if st == 5:
            raise ReferencePanic("color component exceeds the max value")
        dt = np.uint8 if max_value <= 255 else np.uint16
        return Image(width, height, samples=samples.astype(dt).reshape(height, width, 3), max_value=max_value)

    def read_image(self) -> Image:
        data = self.reader.read()
        if isinstance(data, str):
            data = data.encode()
        if self.native:
            return self._read_native(data)
        data = _COMMENT.sub(b"", data)
        tokens = [t for t in _RUST_WS.split(data) if t]
        if not tokens or tokens[0] != b"P3":
            raise PPMFileDoesNotContainRequiredToken(P3_HEADER_TOKEN_NAME)
        hdr = []
        for i, name in ((1, WIDTH_HEADER_TOKEN_NAME), (2, HEIGHT_HEADER_TOKEN_NAME), (3, MAX_VALUE_HEADER_TOKEN_NAME)):
            if len(tokens) <= i:
                raise PPMFileDoesNotContainRequiredToken(name)
            hdr.append(_parse_u16(tokens[i], name))
        width, height, max_value = hdr
        body = tokens[4:]
        if body and max(map(len, body)) <= 5 and all(t.isdigit() for t in body):
            vals = np.array(body, dtype="S5").astype(np.int64)
            if vals.size and int(vals.max()) > 65535:
                raise ParsingOfTokenFailed(COLOR_COMPONENT_VALUE_TOKEN_NAME)
        else:
            vals = np.array([_parse_u16(t, COLOR_COMPONENT_VALUE_TOKEN_NAME) for t in body], dtype=np.int64)
        if vals.size % 3:
            raise IncompletePixelParsed(vals.size % 3)
        if vals.size // 3 != width * height:
            raise MismatchOfSizeBetweenHeaderAndValues()
        if vals.size and int(vals.max()) > max_value:
            # RangeColorFormat::new panics (color.rs:62-65)
            raise ReferencePanic("color component exceeds the max value")
        dt = np.uint8 if max_value <= 255 else np.uint16
        return Image(width, height, samples=vals.astype(dt).reshape(height, width, 3), max_value=max_value)


def parse_ppm_native(data, threads: int = 1):
    """The library's host-side P3 tokenizer (dmmt_ppm_parse, csrc/ppm_parse.hpp: needs no GPU) ->
    (status, detail, width, height, max_value, samples as a uint16 array).  status 0 = OK, 1..5 = DMMT_PPM_*."""
    import ctypes as C

    from . import _ffi as F

    if isinstance(data, str):
        data = data.encode()
    L = F.lib()
    w, h, m = C.c_uint16(), C.c_uint16(), C.c_uint16()
    ptr, n, detail = C.POINTER(C.c_uint16)(), C.c_size_t(), C.c_int()
    st = L.dmmt_ppm_parse(data, len(data), threads, C.byref(w), C.byref(h), C.byref(m), C.byref(ptr), C.byref(n),
                          C.byref(detail))
    if st < 0:
        raise F.DmmtError(st, "dmmt_ppm_parse")
    samples = np.empty(0, np.uint16)
    if st == 0:
        samples = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else samples
        L.dmmt_free(ptr)
    return st, detail.value, w.value, h.value, m.value, samples


def ppm_error_text(status: int, detail: int) -> str:
    import ctypes as C

    from . import _ffi as F

    buf = C.create_string_buffer(256)
    return F.lib().dmmt_ppm_strerror(status, detail, buf, 256).decode()


# ----------------------------------------------------------------------------------- the writer
@dataclass
class JpegTransformationOptions:
    chroma_subsampling_preset: ChromaSubsamplingPreset = ChromaSubsamplingPreset.P420
    bits_per_channel: int = 8
    quantization_table_preset: QuantizationTablePreset = QuantizationTablePreset.Specification

    @classmethod
    def from_arguments(cls, a: "Arguments") -> "JpegTransformationOptions":
        return cls(a.chroma_subsampling_preset, a.bits_per_channel, a.quantization_table_preset)

    def to_c(self) -> Options:
        return Options(self.chroma_subsampling_preset.value, self.bits_per_channel,
                       self.quantization_table_preset.value)


_default_context: Context | None = None


def default_context() -> Context:
    global _default_context
    if _default_context is None:
        _default_context = Context(int(os.environ.get("LOCAL_RANK", "0")) if F.lib().dmmt_device_count() > 1 else 0)
    return _default_context


class JpegImageWriter:
    """JpegImageWriter::new(writer, &image, &options, &threadpool) (jpeg.rs:48-62).  `threadpool`
    is accepted for signature parity; the reference only uses it to fan the DCT out
    (transformer.rs:126-148), which the GPU does on its own."""

    def __init__(self, writer, image: Image, options: JpegTransformationOptions, threadpool=None,
                 context: Context | None = None):
        self.writer, self.image, self.options, self.threadpool = writer, image, options, threadpool
        self.context = context

    def write_image(self) -> None:
        ctx = self.context or default_context()
        im = self.image
        try:
            if im.samples is not None:
                data = ctx.encode(im.samples, im.max_value, self.options.to_c())
            else:
                data = ctx.encode(im.dots, 1, self.options.to_c())
        except F.DmmtError as e:
            if e.code == F.E_SYMBOL:
                raise HuffmanSymbolNotPresentInTranslator() from e
            if e.code == F.E_WRITE:
                raise FailedToWriteImageData() from e
            if e.code in (F.E_RANGE, F.E_INVALID, F.E_SIZE):
                raise ReferencePanic(str(e)) from e
            raise
        self.writer.write(data)
        self.writer.flush()


# ------------------------------------------------------------------------------------------ CLI
@dataclass
class Arguments:
    input_file: str
    output_file: str
    bits_per_channel: int = 8
    chroma_subsampling_preset: ChromaSubsamplingPreset = ChromaSubsamplingPreset.P420
    number_of_threads: int = 1
    quantization_table_preset: QuantizationTablePreset = QuantizationTablePreset.Specification


class CLIParser:
    """cli.rs:12-180: positional input_file output_file; -b/--bits_per_channel {8,16,32} (8);
    -p/--chroma_subsampling_preset {P444,P422,P420} (P420); -t/--threads (available parallelism);
    -q/--quantization_table (Specification).  Usage errors exit with status 2 like clap."""

    def __init__(self):
        p = argparse.ArgumentParser(prog="dmmt-jpeg-encoder")
        p.add_argument("input_file", help="Path to PPM imput file")
        p.add_argument("output_file", help="Path to JPEG output file")
        p.add_argument("-b", "--bits_per_channel", metavar="BITS", default="8", choices=["8", "16", "32"],
                       help="Bits per color channel")
        p.add_argument("-p", "--chroma_subsampling_preset", metavar="PRESET", default="P420",
                       choices=["P444", "P422", "P420"], help="Chroma subsampling preset")
        p.add_argument("-t", "--threads", metavar="THREADS", type=self._usize, default=os.cpu_count() or 1,
                       help="Number of Threads")
        p.add_argument("-q", "--quantization_table", metavar="TABLE", default="Specification",
                       choices=list(QUANTIZATION_TABLE_CLI_NAMES), help="Quantization table preset")
        self.command = p

    @staticmethod
    def _usize(s: str) -> int:
        v = int(s)
        if v < 0:
            raise argparse.ArgumentTypeError("invalid digit found in string")
        return v

    @classmethod
    def default(cls) -> "CLIParser":
        return cls()

    def parse(self, itr) -> Arguments:
        argv = list(itr)[1:]  # args_os() includes the program name
        m = self.command.parse_args(argv)
        return Arguments(m.input_file, m.output_file, int(m.bits_per_channel),
                         ChromaSubsamplingPreset[m.chroma_subsampling_preset], m.threads,
                         QUANTIZATION_TABLE_CLI_NAMES[m.quantization_table])


def convert_ppm_to_jpeg(arguments: Arguments, context: Context | None = None) -> None:
    """lib.rs:59-77: opens the input, creates/truncates the output (before parsing), reads the
    P3 image, writes the JPEG."""
    try:
        fin = open(arguments.input_file, "rb")
    except OSError as e:
        raise UnableToOpenInputFileForReading(arguments.input_file, f"{e.strerror} (os error {e.errno})") from e
    with fin:
        try:
            fout = open(arguments.output_file, "wb")
        except OSError as e:
            raise UnableToOpenOutputFileForWriting(arguments.output_file, f"{e.strerror} (os error {e.errno})") from e
        with fout:
            image = PPMImageReader(fin, native=True, threads=max(1, int(arguments.number_of_threads))).read_image()
            options = JpegTransformationOptions.from_arguments(arguments)
            JpegImageWriter(fout, image, options, None, context).write_image()


def main(argv=None) -> int:
    """main.rs:5-12: prints the outcome; exit status 0 on success AND on conversion failure."""
    arguments = CLIParser.default().parse(sys.argv if argv is None else argv)
    try:
        convert_ppm_to_jpeg(arguments)
        print("Conversion successful")
    except Error as e:
        print(f"Conversion failed because of: {e}", file=sys.stderr)
    return 0
