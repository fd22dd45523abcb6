"""Size of the GPU deflate parse (host emulation of the kernel's logic, tests/emul/deflate_emul.cpp) for
several candidate-distance sets, against zlib level 6 and zlib's Z_RLE, on index planes of the CPU oracle.
CPU only.  python profiles/deflate_candidates.py [H W]"""
import ctypes, os, subprocess, sys, zlib, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vcf_oracle as O

SO = os.path.join(ROOT, "build", "deflate_emul.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", SO,
                       os.path.join(ROOT, "tests", "emul", "deflate_emul.cpp")])
L = ctypes.CDLL(SO)
L.dfl_emul.restype = ctypes.c_longlong
L.dfl_emul.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_int,
                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong)]


def emul(data, row=0, pixel=1, dists=None, model=1):
    data = np.ascontiguousarray(data, np.uint8).ravel()
    cap = data.size + data.size // 100 + 4096
    out = np.empty(cap, np.uint8)
    d = np.asarray(dists if dists is not None else [], np.int32)
    n = L.dfl_emul(data.ctypes.data, data.size, 258, 512, row, pixel, d.ctypes.data, d.size, model, out.ctypes.data, cap, None)
    assert n > 0, n
    raw = out[:n].tobytes()
    assert zlib.decompress(raw, -15) == data.tobytes()
    return n


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
    res = {}
    for kind in ("natural", "noise"):
        img = O.synthetic_frame(H, W, 2, kind)
        for q in (8, 16, 32, 64):
            idx = O.encode_array(img, 8, q)
            row = idx.shape[1] * idx.shape[2]          # bytes per row of the H x W x 3 array the entropy stage gets
            b = idx.tobytes()
            c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
            r = {"shape": list(idx.shape), "row": row, "zlib6": len(zlib.compress(b, 6)), "zlib_rle": len(c.compress(b) + c.flush()),
                 "runs_plain": emul(idx, model=0), "runs": emul(idx), "default": emul(idx, row=row, pixel=3)}
            for name, d in (("d3", [1, 3]), ("d3_6", [1, 3, 6]), ("d3_row", [1, 3, row]), ("d3_row_pm3", [1, 3, row, row - 3, row + 3]),
                            ("d3_6_row_pm3", [1, 3, 6, row, row - 3, row + 3]), ("d3_6_9_row_pm3_pm6", [1, 3, 6, row, row - 3, row + 3, row - 6, row + 6])):
                r[name] = emul(idx, dists=d)
            res[f"{kind}_q{q}"] = r
            print(kind, q, r, flush=True)
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2_deflate_candidates.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
