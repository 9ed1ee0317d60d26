"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, the
host-side mirror of the reference interface (P3 reader, CLI, errors), and the sharded exchange
logic over gloo with world_size 2/3 (simulated shards, no GPU)."""
import ctypes as C
import io
import os
import re
import socket

import numpy as np
import pytest

from conftest import ROOT, synth_image


def test_library_builds_loads_and_exports_every_declared_symbol():
    import dmmt_jpeg_encoder_b200 as D
    from dmmt_jpeg_encoder_b200 import _ffi as F

    header = open(os.path.join(ROOT, "include", "dmmt_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dmmt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    L = F.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/dmmt_cuda.h but not exported"
    assert declared == set(F.SIGNATURES), declared ^ set(F.SIGNATURES)
    assert L.dmmt_strerror(0) == b"ok" and b"no CPU fallback" in L.dmmt_strerror(F.E_NODEVICE)


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import dmmt_jpeg_encoder_b200 as D
    from dmmt_jpeg_encoder_b200 import _ffi as F

    if F.lib().dmmt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(D.DmmtError) as e:
        D.Context(0)
    assert e.value.code == F.E_NODEVICE
    img = D.Image(8, 8, samples=np.zeros((8, 8, 3), np.uint8), max_value=255)
    with pytest.raises(D.DmmtError):
        D.JpegImageWriter(io.BytesIO(), img, D.JpegTransformationOptions()).write_image()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dmmt_jpeg_encoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


# ---------------------------------------------------------------------- P3 reader (ppm.rs tests)
def _read(text):
    from dmmt_jpeg_encoder_b200 import PPMImageReader

    return PPMImageReader(io.BytesIO(text.encode() if isinstance(text, str) else text)).read_image()


def test_ppm_reader_basic_and_comments():
    im = _read("P3\n# a comment\n2 2\n255\n255 0 0  0 255 0\n0 0 255 # trailing\n 10 20 30\n")
    assert (im.width, im.height, im.max_value) == (2, 2, 255)
    assert im.samples.tolist() == [[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [10, 20, 30]]]
    np.testing.assert_array_equal(im.dots, im.samples.astype(np.float32) / np.float32(255))
    # a comment runs through the newline and does not split the token around it (ppm.rs:49-60)
    im = _read("P3 1 1 255 1#x\n2 3 4")
    assert im.samples.tolist() == [[[12, 3, 4]]]


def test_ppm_reader_errors_match_reference_messages():
    import dmmt_jpeg_encoder_b200.reference_api as R

    with pytest.raises(R.PPMFileDoesNotContainRequiredToken) as e:
        _read("P6 1 1 255 0 0 0")
    assert str(e.value) == "Expected token 'P3 Header' not found in PPM file"
    with pytest.raises(R.PPMFileDoesNotContainRequiredToken) as e:
        _read("P3 4")
    assert "Height Header" in str(e.value)
    with pytest.raises(R.ParsingOfTokenFailed) as e:
        _read("P3 x 1 255 0 0 0")
    assert str(e.value) == "Parsing of token 'Width Header' failed"
    with pytest.raises(R.ParsingOfTokenFailed):
        _read("P3 1 1 255 0 0 70000")
    with pytest.raises(R.ParsingOfTokenFailed):
        _read("P3 1 1 255 0 -1 0")
    with pytest.raises(R.IncompletePixelParsed) as e:
        _read("P3 1 1 255 0 0 0 9 9")
    assert str(e.value) == "Incomplete pixel parsed. Expected 3 components, but got 2."
    with pytest.raises(R.MismatchOfSizeBetweenHeaderAndValues) as e:
        _read("P3 2 1 255 0 0 0")
    assert str(e.value) == "Nubmer of pixels do not match the size, provided in header"
    with pytest.raises(R.ReferencePanic):
        _read("P3 1 1 15 0 16 0")
    assert _read("P3 1 1 65535 +5 007 65535").samples.tolist() == [[[5, 7, 65535]]]


def test_ppm_reader_agrees_with_oracle_parser_on_fixtures():
    from conftest import FIXTURES, load_fixture
    from oracle import oracle as O

    for name in FIXTURES:
        text, px, mx = load_fixture(name)
        im = _read(text)
        w, h, m, s = O.parse_ppm(text)
        assert (im.width, im.height, im.max_value) == (w, h, m)
        np.testing.assert_array_equal(im.samples, s)


# ------------------------------------------ native P3 tokenizer (csrc/ppm_parse.hpp, dmmt_ppm_parse)
def _native(text, threads=1):
    import dmmt_jpeg_encoder_b200.reference_api as R

    return R.parse_ppm_native(text, threads)


def _expect_from_python_reader(text):
    """(status, detail, samples) the native tokenizer must produce, derived from the Python mirror of ppm.rs."""
    import dmmt_jpeg_encoder_b200.reference_api as R

    names = [R.P3_HEADER_TOKEN_NAME, R.WIDTH_HEADER_TOKEN_NAME, R.HEIGHT_HEADER_TOKEN_NAME,
             R.MAX_VALUE_HEADER_TOKEN_NAME, R.COLOR_COMPONENT_VALUE_TOKEN_NAME]
    try:
        im = _read(text)
    except R.PPMFileDoesNotContainRequiredToken as e:
        return 1, [n in str(e) for n in names].index(True), None
    except R.ParsingOfTokenFailed as e:
        return 2, [n in str(e) for n in names].index(True), None
    except R.IncompletePixelParsed as e:
        return 3, int(str(e).split("got ")[1].rstrip(".")), None
    except R.MismatchOfSizeBetweenHeaderAndValues:
        return 4, 0, None
    except R.ReferencePanic:
        return 5, 0, None
    return 0, 0, np.asarray(im.samples).reshape(-1).astype(np.uint16)


def test_native_ppm_tokenizer_known_cases_and_error_texts():
    import dmmt_jpeg_encoder_b200.reference_api as R

    st, det, w, h, m, s = _native("P3\n# a comment\n2 2\n255\n255 0 0  0 255 0\n0 0 255 # trailing\n 10 20 30\n")
    assert (st, w, h, m) == (0, 2, 2, 255) and s.tolist() == [255, 0, 0, 0, 255, 0, 0, 0, 255, 10, 20, 30]
    assert _native("P3 1 1 255 1#x\n2 3 4")[5].tolist() == [12, 3, 4]          # comment inside a token
    assert _native("P3 1 1 65535 +5 007 65535")[5].tolist() == [5, 7, 65535]   # '+', leading zeros
    assert _native("P3 1 1 65535 0000000012 3 4")[5].tolist() == [12, 3, 4]    # more than five digits
    cases = {"P6 1 1 255 0 0 0": "Expected token 'P3 Header' not found in PPM file",
             "": "Expected token 'P3 Header' not found in PPM file",
             "P3 4": "Expected token 'Height Header' not found in PPM file",
             "P3 x 1 255 0 0 0": "Parsing of token 'Width Header' failed",
             "P3 1 1 255 0 0 70000": "Parsing of token 'Color Component Value' failed",
             "P3 1 1 255 0 -1 0": "Parsing of token 'Color Component Value' failed",
             "P3 1 1 255 0 1\x0b2 0": "Parsing of token 'Color Component Value' failed",   # \x0B is no whitespace
             "P3 1 1 255 0 + 0": "Parsing of token 'Color Component Value' failed",
             "P3 1 1 255 0 0 0 9 9": "Incomplete pixel parsed. Expected 3 components, but got 2.",
             "P3 2 1 255 0 0 0": "Nubmer of pixels do not match the size, provided in header"}
    for text, msg in cases.items():
        st, det = _native(text)[:2]
        assert st > 0 and R.ppm_error_text(st, det) == msg, (text, st, det)
    assert _native("P3 1 1 15 0 16 0")[0] == 5   # the reference panics (color.rs:62-65)


def test_native_ppm_tokenizer_fuzz_against_the_python_reader():
    """Random token soups (all separators, comments inside and between tokens, '+', long and overflowing numbers,
    stray bytes) of lengths that cross the 64-byte blocks of the fast path: same samples / same error."""
    rng = np.random.default_rng(20261018)
    seps = [" ", "\n", "\t", "\r", "\x0c", "  ", " \n", "\r\n"]
    for case in range(400):
        n = int(rng.integers(0, 80)) * 3
        w = max(1, n // 3)
        mx = int(rng.choice([255, 65535, 1023]))
        toks = [str(int(v)) for v in rng.integers(0, mx + 1, n)]
        flavour = case % 8
        for i in range(len(toks)):
            r = rng.random()
            if flavour >= 2 and r < 0.03:
                toks[i] = "+" + toks[i]
            elif flavour >= 2 and r < 0.06:
                toks[i] = "0" * int(rng.integers(1, 7)) + toks[i]
            elif flavour >= 4 and r < 0.09 and len(toks[i]) > 1:
                toks[i] = toks[i][0] + "#c " + str(i) + "\n" + toks[i][1:]      # comment inside the token
            elif flavour >= 6 and r < 0.10:
                toks[i] = rng.choice(["70000", "65536", "1x", "-1", "1\x0b", "+", "99999999999"])
        text = f"P3 {w} 1 {mx}" + rng.choice(seps)
        for t in toks:
            text += t + (rng.choice(seps) if rng.random() < 0.9 else " # note\n")
        if flavour == 1 and n:
            text = text.rstrip()                                                   # number ends the buffer
        if flavour == 3:
            text += " 7"                                                           # incomplete pixel / mismatch
        est, edet, esam = _expect_from_python_reader(text)
        for threads in (1, 3):
            st, det, gw, gh, gm, s = _native(text, threads)
            assert (st, det) == (est, edet), (case, text[:200], st, det, est, edet)
            if st == 0:
                np.testing.assert_array_equal(s, esam)


def test_python_reader_native_mode_equals_pure_python_mode():
    """PPMImageReader(native=True) (what convert_ppm_to_jpeg uses) raises the same errors with the same texts and
    returns the same Image as the pure-Python mirror."""
    from conftest import FIXTURES, load_fixture
    from dmmt_jpeg_encoder_b200 import PPMImageReader
    import dmmt_jpeg_encoder_b200.reference_api as R

    def both(text):
        out = []
        for native in (False, True):
            try:
                im = PPMImageReader(io.BytesIO(text.encode()), native=native, threads=2).read_image()
                out.append(("ok", im.width, im.height, im.max_value, im.samples.dtype, im.samples.tolist()))
            except (R.Error, R.ReferencePanic) as e:
                out.append((type(e).__name__, str(e)))
        return out

    texts = ["P6 1 1 255 0 0 0", "P3 4", "P3 x 1 255 0 0 0", "P3 1 1 255 0 0 70000", "P3 1 1 255 0 0 0 9 9",
             "P3 2 1 255 0 0 0", "P3 1 1 15 0 16 0", "P3 1 1 65535 +5 007 65535", "P3 1 1 255 1#x\n2 3 4", ""]
    texts += [load_fixture(name)[0] for name in FIXTURES]
    for text in texts:
        a, b = both(text if isinstance(text, str) else text.decode())
        assert a == b, (text[:60], a[:2], b[:2])


def test_native_ppm_tokenizer_large_file_threads_agree():
    """A 2.4 MB comment-free file takes the multi-threaded path: every thread count gives the same samples; one bad
    token anywhere fails the parse; a '#' anywhere falls back to one thread and still agrees."""
    rng = np.random.default_rng(7)
    w, h = 640, 320
    px = rng.integers(0, 65536, w * h * 3)
    body = "".join(f"{v}{' ' if (i + 1) % 17 else chr(10)}" for i, v in enumerate(px.tolist()))
    text = f"P3\n{w} {h}\n65535\n" + body
    assert len(text) > (1 << 20)
    ref = None
    for threads in (1, 2, 5, 8, 64, 1000):
        st, det, gw, gh, gm, s = _native(text, threads)
        assert (st, gw, gh, gm) == (0, w, h, 65535)
        np.testing.assert_array_equal(s, px.astype(np.uint16))
    cut = len(text) * 3 // 4
    cut = text.index(" ", cut)
    bad = text[:cut] + " 65536" + text[cut:]
    assert _native(bad, 8)[:2] == (2, 4) and _native(bad, 1)[:2] == (2, 4)
    commented = text[:cut] + " # a comment 1 2 3\n" + text[cut:]
    np.testing.assert_array_equal(_native(commented, 8)[5], px.astype(np.uint16))


# ------------------------------------------------------------------------------------------ CLI
def test_cli_defaults_and_aliases():
    from dmmt_jpeg_encoder_b200 import ChromaSubsamplingPreset, CLIParser, QuantizationTablePreset

    a = CLIParser.default().parse(["prog", "in.ppm", "out.jpg"])
    assert (a.input_file, a.output_file, a.bits_per_channel) == ("in.ppm", "out.jpg", 8)
    assert a.chroma_subsampling_preset is ChromaSubsamplingPreset.P420
    assert a.quantization_table_preset is QuantizationTablePreset.Specification
    assert a.number_of_threads == (os.cpu_count() or 1)
    a = CLIParser().parse(["prog", "-b", "16", "-p", "P422", "-t", "3", "-q", "6", "a", "b"])
    assert (a.bits_per_channel, a.number_of_threads) == (16, 3)
    assert a.chroma_subsampling_preset is ChromaSubsamplingPreset.P422
    assert a.quantization_table_preset is QuantizationTablePreset.DCTunePerceptualOptimization
    for spelling, want in [("Spec", 0), ("Default", 0), ("Flat", 1), ("MSSIM-Kodak-Tuned", 2), ("4", 3),
                           ("A-visual-detection-model", 5), ("8", 6)]:
        a = CLIParser().parse(["prog", "--quantization_table", spelling, "a", "b"])
        assert a.quantization_table_preset.value == want
    for bad in (["prog", "a"], ["prog", "-b", "12", "a", "b"], ["prog", "-q", "3", "a", "b"],
                ["prog", "-p", "p420", "a", "b"]):
        with pytest.raises(SystemExit) as e:
            CLIParser().parse(bad)
        assert e.value.code == 2


def test_convert_reports_unopenable_input(tmp_path, capsys):
    import dmmt_jpeg_encoder_b200.reference_api as R

    missing = str(tmp_path / "nope.ppm")
    with pytest.raises(R.UnableToOpenInputFileForReading) as e:
        R.convert_ppm_to_jpeg(R.Arguments(missing, str(tmp_path / "o.jpg")))
    assert str(e.value) == f"Unable to open input file '{missing}' for reading: No such file or directory (os error 2)"
    assert R.main(["prog", missing, str(tmp_path / "o.jpg")]) == 0          # main.rs: exit 0 either way
    assert "Conversion failed because of: Unable to open input file" in capsys.readouterr().err


# ------------------------------------------------------------------------------ C++ front-end
def _cli():
    from dmmt_jpeg_encoder_b200 import build as B

    B.build()
    return B.CLI


def test_cpp_cli_usage_errors_and_messages(tmp_path):
    """dmmt-jpeg-encoder (csrc/cli_main.cpp + dmmt_host.hpp) mirrors src/main.rs + src/cli.rs: clap-style
    usage errors exit 2; conversion failures print the reference's Display text and exit 0."""
    import subprocess

    cli = _cli()
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2 and "<input_file>" in r.stderr
    for bad in (["-b", "12", "a", "b"], ["-q", "3", "a", "b"], ["-p", "p420", "a", "b"], ["-t", "x", "a", "b"], ["a", "b", "c"]):
        assert subprocess.run([cli, *bad], capture_output=True).returncode == 2, bad
    missing = str(tmp_path / "nope.ppm")
    r = subprocess.run([cli, missing, str(tmp_path / "o.jpg")], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stderr.strip() == (f"Conversion failed because of: Unable to open input file '{missing}' for reading: "
                                "No such file or directory (os error 2)")
    # the P3 parse (host I/O) runs before any CUDA call, so its errors are testable without a GPU
    cases = {"P6 1 1 255 0 0 0": "Expected token 'P3 Header' not found in PPM file",
             "P3 x 1 255 0 0 0": "Parsing of token 'Width Header' failed",
             "P3 1 1 255 0 0 70000": "Parsing of token 'Color Component Value' failed",
             "P3 1 1 255 0 0 0 9 9": "Incomplete pixel parsed. Expected 3 components, but got 2.",
             "P3 2 1 255 0 0 0": "Nubmer of pixels do not match the size, provided in header"}
    for text, msg in cases.items():
        src = tmp_path / "in.ppm"
        src.write_text(text)
        r = subprocess.run([cli, str(src), str(tmp_path / "o.jpg")], capture_output=True, text=True)
        assert r.returncode == 0 and r.stderr.strip() == "Conversion failed because of: " + msg, (text, r.stderr)
        assert (tmp_path / "o.jpg").exists()          # created/truncated before parsing (lib.rs:60-61)
    r = subprocess.run([cli, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--chroma_subsampling_preset" in r.stdout


# --------------------------------------------------------------------------- sharded exchange
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,preset,w,h", [(2, 2, 40, 70), (2, 0, 24, 40), (3, 1, 33, 50), (2, 2, 8, 40)])
def test_sharded_exchange_logic_over_gloo(world, preset, w, h):
    """world_size > 1 over gloo: simulated shards + the real encode_sharded driver must stitch the
    oracle's whole-file output (DC seeds, global tables, bit offsets, tail hand-over, byte gather)."""
    import torch.multiprocessing as mp

    import _shard_sim

    px = synth_image("photo" if w > 8 else "grad", w, h, 9) if w > 8 else np.full((h, w, 3), 90, np.uint8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_sim.worker, args=(r, world, port, px, preset, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, whole = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out == whole


@pytest.mark.parametrize("phase", ["transform", "histogram", "tables", "pack", "stuff"])
def test_sharded_failure_reaches_every_rank(phase):
    """A shard that fails in any phase must not strand the other ranks in a collective: every rank raises the same
    DmmtError (here E_SYMBOL on rank 1 of 3)."""
    import torch.multiprocessing as mp

    import _shard_sim
    from dmmt_jpeg_encoder_b200 import _ffi as F

    px = synth_image("photo", 40, 70, 3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_sim.failing_worker, args=(r, 3, port, px, 2, q, 1, phase, F.E_SYMBOL, 9))
             for r in range(3)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120)[:3] for _ in range(3))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == [(r, "error", F.E_SYMBOL) for r in range(3)]


def test_sharded_overflow_is_retried_on_every_rank():
    """DMMT_E_OVERFLOW on one shard: ALL ranks grow their shard and the phases run once more (the reference encodes
    any input); a second overflow is reported on every rank."""
    import torch.multiprocessing as mp

    import _shard_sim
    from dmmt_jpeg_encoder_b200 import _ffi as F

    px = synth_image("photo", 40, 70, 4)
    for times, want in ((1, "ok"), (2, "error")):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_shard_sim.failing_worker, args=(r, 2, port, px, 2, q, 1, "tables", F.E_OVERFLOW, times))
                 for r in range(2)]
        for p in procs:
            p.start()
        got = sorted(q.get(timeout=120) for _ in range(2))
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        assert [g[1] for g in got] == [want, want]
        assert [g[4] for g in got] == [1, 1]                # grow() once on each rank
        if want == "ok":
            assert got[0][2] == got[0][3]                   # rank 0 holds the whole file
        else:
            assert [g[2] for g in got] == [F.E_OVERFLOW, F.E_OVERFLOW]


def test_batch_front_end_usage_errors():
    """dmmt-jpeg-batch keeps the options of the reference CLI; usage errors exit with status 2 like clap."""
    import subprocess

    from dmmt_jpeg_encoder_b200 import build as B

    B.build()
    for bad in ([], ["outdir"], ["-p", "P411", "outdir", "a.ppm"], ["--nope", "outdir", "a.ppm"], ["-b", "12", "o", "a.ppm"]):
        r = subprocess.run([B.CLI_BATCH, *bad], capture_output=True, text=True)
        assert r.returncode == 2, (bad, r.stderr)
        assert "error:" in r.stderr


def test_rust_shim_matches_the_header():
    """rust/ cannot be compiled here (no rustc): its #[repr(C)] structs, extern "C" signatures and constants are
    compared with include/dmmt_cuda.h by rust/check_layout.py instead."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "rust", "check_layout.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


def test_failed_collective_is_dmmt_e_nccl():
    """A collective of the sharded drivers that fails (here: no process group) surfaces as DmmtError(DMMT_E_NCCL),
    the code include/dmmt_cuda.h reserves for the exchanges of the sharded path."""
    import torch

    from dmmt_jpeg_encoder_b200 import _ffi as F
    from dmmt_jpeg_encoder_b200 import sharded as S

    with pytest.raises(F.DmmtError) as e:
        S.dist.all_reduce(torch.zeros(4))
    assert e.value.code == F.E_NCCL and "all_reduce" in str(e.value)
    assert S.dist.ReduceOp.SUM is not None        # classes and constants pass through untouched


def test_shard_row_partition():
    from dmmt_jpeg_encoder_b200 import sharded as S
    from dmmt_jpeg_encoder_b200.encoder import Options

    assert S.mcu_rows_total(32768, Options(2)) == 2048 and S.mcu_rows_total(17, Options(2)) == 2
    assert S.mcu_rows_total(17, Options(1)) == 3
    rows = [S.shard_rows(2048, 8, r) for r in range(8)]
    assert rows[0] == (0, 256) and rows[-1] == (1792, 2048)
    assert all(rows[i][1] == rows[i + 1][0] for i in range(7))
    assert S.pixel_row_range(1080, Options(2), 60, 68) == (960, 1080)


def test_mailbox_and_debug_entry_points_validate_their_arguments():
    """the exchange / test-hook entry points added in round 2 reject bad arguments before touching a device"""
    import ctypes as C

    from dmmt_jpeg_encoder_b200 import _ffi as F

    L = F.lib()
    assert L.dmmt_mailbox_bytes(0) == 0 and L.dmmt_mailbox_bytes(65) == 0
    one, eight = L.dmmt_mailbox_bytes(1), L.dmmt_mailbox_bytes(8)
    assert one > 0 and eight == 8 * one                      # slots x world rows of payload + flag
    got = C.c_size_t(0)
    buf = (C.c_uint8 * 16)()
    assert L.dmmt_debug_stuff(None, buf, 16, 0, buf, 16, C.byref(got)) == F.E_INVALID
    boxes = (C.c_void_p * 2)()
    assert L.dmmt_shard_launch_post(None, boxes, 0, 2, 0, 1, buf, 1) == F.E_INVALID
    assert L.dmmt_shard_launch_collect(None, buf, 2, 0, 1, 0, 1, buf) == F.E_INVALID
