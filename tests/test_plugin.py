"""The drop-in spatial-transform module (vcf_b200/plugin/2D-DCT-B200.py) driven through a
CoDec chain, exactly as the reference drives src/2D-DCT.py: CLI, flags, side files."""
import io
import os
import struct
import subprocess
import sys

import cv2
import numpy as np
import pytest

from oracle import vcf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "vcf_stub")
PLUGIN_DIR = os.path.join(ROOT, "vcf_b200", "plugin")
PLUGIN = os.path.join(PLUGIN_DIR, "2D-DCT-B200.py")
REF_SRC = "/root/reference/src"


def _run(cwd, script, *argv, extra_path=()):
    if cwd == STUB and script != "-c" and "-t" not in argv:
        argv = (*argv, "-t", "chain_stub")          # the stand-in chain lives in one module
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([PLUGIN_DIR, ROOT, *extra_path, env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return subprocess.run([sys.executable, script, *argv], cwd=cwd, env=env, capture_output=True, text=True)


def _write_png(fn, img):
    assert cv2.imwrite(fn, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))


def _read_png(fn):
    return cv2.cvtColor(cv2.imread(fn, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)


@pytest.mark.gpu
@pytest.mark.parametrize("flags,kw", [
    ([], dict(B=8, q=32)),
    (["-B", "16", "-q", "8"], dict(B=16, q=8)),
    (["-q", "12", "-x"], dict(B=8, q=12, disable_subbands=True)),
    (["-q", "4", "-p"], dict(B=8, q=4, perceptual=True)),
    (["-t", "chain_stub_alt"], dict(B=8, q=32)),     # -t only changes the base class (src/2D-DCT.py:22-23)
])
def test_cli_encode_decode_matches_reference_semantics(flags, kw):
    img = O.synthetic_frame(136, 200, 31, "natural")       # 136 = 17*8: padding for B=16
    _write_png("/tmp/original.png", img)
    for f in ("/tmp/encoded.npz", "/tmp/encoded_shape.bin", "/tmp/decoded.png"):
        if os.path.exists(f):
            os.remove(f)
    r = _run(STUB, PLUGIN, "encode", *flags)
    assert r.returncode == 0, r.stderr[-2000:]
    assert struct.unpack("iii", open("/tmp/encoded_shape.bin", "rb").read()) == img.shape
    idx = np.load("/tmp/encoded.npz")["a"]
    assert np.array_equal(idx, O.encode_array(img, **kw))
    r = _run(STUB, PLUGIN, "decode", *flags)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(_read_png("/tmp/decoded.png"), O.decode_array(idx, img.shape, **kw))


@pytest.mark.gpu
def test_cli_post_filter_gets_unclipped_float():
    img = O.synthetic_frame(64, 128, 32, "natural")
    _write_png("/tmp/original.png", img)
    assert _run(STUB, PLUGIN, "encode", "-q", "16").returncode == 0
    r = _run(STUB, PLUGIN, "decode", "-q", "16", "-f", "half_filter")
    assert r.returncode == 0, r.stderr[-2000:]
    idx = np.load("/tmp/encoded.npz")["a"]
    yf = O.decode_array(idx, img.shape, 8, 16, return_float=True)
    want = np.clip(yf * 0.5 + 300.25, 0, 255).astype(np.uint8)
    assert np.array_equal(_read_png("/tmp/decoded.png"), want)


@pytest.mark.gpu
def test_iii_style_per_frame_loop_and_block_size_optimiser():
    frames = [O.synthetic_frame(64, 128, 60 + i, "natural") for i in range(3)]
    for i, f in enumerate(frames):
        _write_png("/tmp/original_%04d.png" % i, f)
    r = _run(STUB, "III_stub.py", "encode", "-N", "3", "-q", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    r = _run(STUB, "III_stub.py", "decode", "-N", "3", "-q", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    for i, f in enumerate(frames):
        idx = np.load("/tmp/encoded_%04d.npz" % i)["a"]
        assert np.array_equal(idx, O.encode_array(f, 8, 16))
        assert np.array_equal(_read_png("/tmp/decoded_%04d.png" % i), O.decode_array(idx, f.shape, 8, 16))
    # -L: J = rate + Lambda*RMSE over 2**i, i = 1..7 (src/2D-DCT.py:533-579; 128 does not divide the frame)
    img = frames[0]
    _write_png("/tmp/original.png", img)
    r = _run(STUB, PLUGIN, "-g", "encode", "-L", "50.0", "-q", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    best, bestJ = None, 1e18
    for B in (2, 4, 8, 16, 32, 64):
        k, y, rm = O.optimize_block_size_point(img, B, 16)     # the reference's loop body as it runs, :538-574
        b = io.BytesIO()
        np.savez_compressed(file=b, a=k)
        J = len(b.getvalue()) + 50.0 * rm
        if J < bestJ:
            best, bestJ = B, J
    assert f"optimal block_size={best}" in r.stderr, r.stderr[-1500:]
    assert np.array_equal(np.load("/tmp/encoded.npz")["a"], O.encode_array(img, best, 16))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["L_q4_bright", "L_q32"])
def test_block_size_search_reproduces_the_references_log(name):
    """The J the plugin logs per block size against the J the UNMODIFIED reference logged for the same image and
    flags (tests/golden/ref_flow_L_*.npz, oracle/make_golden.py), the chosen size and the code-stream."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_flow_{name}.npz"))
    _write_png("/tmp/original.png", g["img"])
    flags = [str(f) for f in g["flags"] if str(f) != "-g"]
    r = _run(STUB, PLUGIN, "-g", "encode", *flags)
    assert r.returncode == 0, r.stderr[-2000:]
    got = {}
    for line in r.stderr.splitlines():
        if "J=" in line and "block_size=" in line:
            got[int(line.split("block_size=")[1].split()[0])] = float(line.split("J=")[1].split()[0])
    assert sorted(got) == [int(b) for b in g["L_block_sizes"]]
    for B, J_ref in zip(g["L_block_sizes"], g["L_J"]):
        assert abs(got[int(B)] - float(J_ref)) <= 2e-6 * float(J_ref), (int(B), got[int(B)], float(J_ref))
    assert f"optimal block_size={int(g['L_chosen'])}" in r.stderr
    assert np.array_equal(np.load("/tmp/encoded.npz")["a"], g["idx"])


def test_plugin_registers_reference_flags_without_gpu():
    """No GPU needed: import-time behaviour, flag names / dests and class chain."""
    r = _run(STUB, PLUGIN, "encode", "-h")
    assert r.returncode == 0, r.stderr[-1500:]
    for flag in ("--block_size_DCT", "--color_transform", "--perceptual_quantization", "--disable_subbands",
                 "--Lambda"):
        assert flag in r.stdout
    r = _run(STUB, PLUGIN, "decode", "-h")
    for flag in ("--block_size_DCT", "--color_transform", "--perceptual_quantization", "--disable_subbands"):
        assert flag in r.stdout
    code = ("import importlib, sys; sys.argv=['x','decode','-B','16','-q','8','-t','chain_stub_alt'];"
            "m = importlib.import_module('2D-DCT-B200'); import parser; c = m.CoDec(parser.parser.parse_known_args()[0]);"
            "print([k.__module__ for k in type(c).__mro__][:3], c.block_size, c.QSS, c.offset)")
    r = _run(STUB, "-c", code)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "'2D-DCT-B200', 'chain_stub_alt', 'chain_stub'" in r.stdout
    assert r.stdout.strip().endswith("16 8 128")


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference only present in the build container")
def test_plugin_against_the_real_reference_chain_cpu():
    """With the reference's own main/parser/YCoCg/deadzone/no_filter/z_lib modules (and the
    shadow packages for its four external imports): same MRO as src/2D-DCT.py, and the
    arithmetic refuses to run without a GPU instead of falling back."""
    shims = os.path.join(ROOT, "oracle", "shims")
    code = ("import importlib, sys; sys.argv=['x','encode','-c','z_lib'];"
            "m = importlib.import_module('2D-DCT-B200'); import parser; c = m.CoDec(parser.parser.parse_known_args()[0]);"
            "print([k.__module__ for k in type(c).__mro__][:6])")
    r = _run(REF_SRC, "-c", code, extra_path=(shims,))
    assert r.returncode == 0, r.stderr[-1500:]
    assert "'2D-DCT-B200', 'YCoCg', 'deadzone', 'no_filter', 'z_lib', 'entropy_image_coding'" in r.stdout
    from vcf_b200 import _lib
    if _lib.lib().vcfb_device_count() == 0:
        img = O.synthetic_frame(16, 16, 1, "noise")
        _write_png("/tmp/original.png", img)
        r = _run(REF_SRC, PLUGIN, "encode", "-c", "z_lib", extra_path=(shims,))
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_with_the_gpu_entropy_stage():
    """-c z_lib-B200 (vcf_b200/plugin/z_lib-B200.py, the drop-in of src/z_lib.py): the code-stream is
    an .npz whose deflate stream came from the GPU; np.load -- what src/z_lib.py:25-29 calls -- and the
    decoder read it."""
    img = O.synthetic_frame(272, 400, 33, "natural")
    _write_png("/tmp/original.png", img)
    for f in ("/tmp/encoded.npz", "/tmp/encoded_shape.bin", "/tmp/decoded.png"):
        if os.path.exists(f):
            os.remove(f)
    r = _run(STUB, PLUGIN, "encode", "-q", "16", "-c", "z_lib-B200")
    assert r.returncode == 0, r.stderr[-2000:]
    idx = np.load("/tmp/encoded.npz")["a"]
    assert idx.dtype == np.uint8 and np.array_equal(idx, O.encode_array(img, B=8, q=16))
    import zipfile
    info = zipfile.ZipFile("/tmp/encoded.npz").infolist()
    assert [i.filename for i in info] == ["a.npy"] and info[0].compress_type == zipfile.ZIP_DEFLATED
    assert info[0].compress_size < idx.size // 4
    r = _run(STUB, PLUGIN, "decode", "-q", "16", "-c", "z_lib-B200")
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(_read_png("/tmp/decoded.png"), O.decode_array(idx, img.shape, B=8, q=16))
    # and the stock entropy stage of the chain reads the same file
    r = _run(STUB, PLUGIN, "decode", "-q", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(_read_png("/tmp/decoded.png"), O.decode_array(idx, img.shape, B=8, q=16))


@pytest.mark.gpu
def test_cli_with_the_gpu_tiff_stage():
    """-c TIFF-B200 (vcf_b200/plugin/TIFF-B200.py, the drop-in of the chain's default src/TIFF.py):
    the code-stream is a deflate TIFF whose strip came from the GPU; libtiff reads it."""
    img = O.synthetic_frame(272, 400, 34, "natural")
    _write_png("/tmp/original.png", img)
    for f in ("/tmp/encoded.tif", "/tmp/encoded_shape.bin", "/tmp/decoded.png"):
        if os.path.exists(f):
            os.remove(f)
    r = _run(STUB, PLUGIN, "encode", "-q", "16", "-c", "TIFF-B200")
    assert r.returncode == 0, r.stderr[-2000:]
    idx = cv2.cvtColor(cv2.imread("/tmp/encoded.tif", cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)
    assert idx.dtype == np.uint8 and np.array_equal(idx, O.encode_array(img, B=8, q=16))
    assert os.path.getsize("/tmp/encoded.tif") < idx.size // 4
    r = _run(STUB, PLUGIN, "decode", "-q", "16", "-c", "TIFF-B200")
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(_read_png("/tmp/decoded.png"), O.decode_array(idx, img.shape, B=8, q=16))


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference only present in the build container")
def test_entropy_plugin_in_the_real_reference_chain_cpu():
    """-c z_lib-B200 inside the reference's own chain: the module takes z_lib's place in the MRO
    (src/no_filter.py:21) and refuses to compress without a GPU."""
    shims = os.path.join(ROOT, "oracle", "shims")
    code = ("import importlib, sys; sys.argv=['x','encode','-c','z_lib-B200'];"
            "m = importlib.import_module('2D-DCT-B200'); import parser; c = m.CoDec(parser.parser.parse_known_args()[0]);"
            "print([k.__module__ for k in type(c).__mro__][:6], c.file_extension);"
            "import numpy as np\n"
            "try: c.compress(np.zeros((8, 8, 3), np.uint8))\n"
            "except Exception as e: print('refused:', e)")
    r = _run(REF_SRC, "-c", code, extra_path=(shims,))
    assert r.returncode == 0, r.stderr[-1500:]
    assert "'2D-DCT-B200', 'YCoCg', 'deadzone', 'no_filter', 'z_lib-B200', 'entropy_image_coding'" in r.stdout
    assert ".npz" in r.stdout
    from vcf_b200 import _lib
    if _lib.lib().vcfb_device_count() == 0:
        assert "refused:" in r.stdout and "no CPU fallback" in r.stdout
    r = _run(REF_SRC, "-c", code.replace("z_lib-B200", "TIFF-B200"), extra_path=(shims,))
    assert r.returncode == 0, r.stderr[-1500:]
    assert "'2D-DCT-B200', 'YCoCg', 'deadzone', 'no_filter', 'TIFF-B200', 'entropy_image_coding'" in r.stdout
    assert ".tif" in r.stdout
    if _lib.lib().vcfb_device_count() == 0:
        assert "refused:" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_batched_iii_driver_equals_per_frame_loop():
    """vcf_b200/plugin/III-B200.py: the whole sequence as one GPU batch must write the same
    files as the per-frame loop of src/III.py:132-144 (and the intended :96-104)."""
    n = 5
    frames = [O.synthetic_frame(72, 128, 80 + i, "natural" if i % 2 else "noise") for i in range(n)]
    for i, f in enumerate(frames):
        _write_png("/tmp/original_%04d.png" % i, f)
        for ext in (".npz", "_shape.bin"):
            if os.path.exists("/tmp/encoded_%04d%s" % (i, ext)):
                os.remove("/tmp/encoded_%04d%s" % (i, ext))
    iii = os.path.join(PLUGIN_DIR, "III-B200.py")
    # with the chain's own entropy stage, and with the GPU one (indices stay in HBM, -c z_lib-B200)
    for entropy_flags in ((), ("-c", "z_lib-B200")):
        for i in range(n):
            for ext in (".npz", "_shape.bin"):
                if os.path.exists("/tmp/encoded_%04d%s" % (i, ext)):
                    os.remove("/tmp/encoded_%04d%s" % (i, ext))
        r = _run(STUB, iii, "encode", "-N", str(n), "-q", "16", *entropy_flags)
        assert r.returncode == 0, r.stderr[-2000:]
        r = _run(STUB, iii, "decode", "-N", str(n), "-q", "16", *entropy_flags)
        assert r.returncode == 0, r.stderr[-2000:]
        for i, f in enumerate(frames):
            assert struct.unpack("iii", open("/tmp/encoded_%04d_shape.bin" % i, "rb").read()) == f.shape
            idx = np.load("/tmp/encoded_%04d.npz" % i)["a"]
            assert np.array_equal(idx, O.encode_array(f, 8, 16))
            assert np.array_equal(_read_png("/tmp/decoded_%04d.png" % i), O.decode_array(idx, f.shape, 8, 16))


@pytest.mark.gpu
def test_in_memory_proxy_equals_file_round_trip():
    """``encode_decode_array`` (no temporary PNGs) gives the hybrid codec the
    same reconstruction, size and payload files as ``encode_fn`` + ``decode_fn`` through files
    (the body of src/IPP_DCT.py:595-626)."""
    img = O.synthetic_frame(72, 136, 90, "natural")
    _write_png("/tmp/proxy_in.png", img)
    code = (
        "import importlib, sys, numpy as np, cv2; sys.argv=['x','encode','-q','12','-t','chain_stub'];"
        "m = importlib.import_module('2D-DCT-B200'); import parser; c = m.CoDec(parser.parser.parse_known_args()[0]);"
        "img = cv2.cvtColor(cv2.imread('/tmp/proxy_in.png'), cv2.COLOR_BGR2RGB);"
        "size_f = c.encode_fn('/tmp/proxy_in.png', '/tmp/proxy_file'); c.decode_fn('/tmp/proxy_file', '/tmp/proxy_file_rec.png');"
        "rec_f = cv2.cvtColor(cv2.imread('/tmp/proxy_file_rec.png'), cv2.COLOR_BGR2RGB);"
        "rec_m, size_m = c.encode_decode_array(img, '/tmp/proxy_mem');"
        "a = np.load('/tmp/proxy_file.npz')['a']; b = np.load('/tmp/proxy_mem.npz')['a'];"
        "print('SAME', bool(np.array_equal(rec_f, rec_m)), size_f == size_m, bool(np.array_equal(a, b)),"
        " open('/tmp/proxy_file_shape.bin','rb').read() == open('/tmp/proxy_mem_shape.bin','rb').read())")
    r = _run(STUB, "-c", code)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SAME True True True True" in r.stdout, r.stdout[-500:]
    rec = O.decode_array(O.encode_array(img, 8, 12), img.shape, 8, 12)
    assert np.array_equal(_read_png("/tmp/proxy_file_rec.png"), rec)
