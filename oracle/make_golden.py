#!/usr/bin/env python
"""Generate tests/golden/ref_flow_*.npz by running the UNMODIFIED reference.

Runs ``/root/reference/src/{2D-DCT,YCoCg,YCrCb}.py encode|decode`` and
``RDE.py`` as sub-processes (cwd = the reference's src/, so its own
``main.py`` / ``parser.py`` / class chain are used untouched) with
``oracle/shims`` on PYTHONPATH standing in for the four un-vendored arithmetic
packages and for skimage.io, and ``-c z_lib`` as the entropy codec (the only one
whose dependencies exist offline; src/z_lib.py).  What is recorded per case:
the input image, the index array found inside the reference's code-stream, the
``_shape.bin`` side file and the decoded PNG.  ``tests/test_reference_flow.py``
then checks ``oracle.vcf_oracle.encode_array/decode_array`` (a restatement that
never touches these scripts) against them -- this pins the oracle's control
flow, padding, dtype chain, bias/wrap and clipping to the reference's.

Only runs in the build container (needs /root/reference); the vectors travel.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from oracle import vcf_oracle as O  # noqa: E402

CASES = [
    # name, H, W, kind, seed, script, flags
    ("default_96x80", 96, 80, "natural", 11, "2D-DCT.py", []),
    ("pad_53x37_q8", 53, 37, "noise", 12, "2D-DCT.py", ["-q", "8"]),
    ("b16_q16", 64, 64, "natural", 13, "2D-DCT.py", ["-B", "16", "-q", "16"]),
    ("b4_q12", 64, 96, "noise", 14, "2D-DCT.py", ["-B", "4", "-q", "12"]),
    ("b32_q64", 96, 64, "natural", 15, "2D-DCT.py", ["-B", "32", "-q", "64"]),
    ("percep_q4", 64, 64, "natural", 16, "2D-DCT.py", ["-q", "4", "-p"]),
    ("nosub", 48, 64, "natural", 17, "2D-DCT.py", ["-x"]),
    ("wrap_b16_q4_pad", 40, 40, "noise", 18, "2D-DCT.py", ["-B", "16", "-q", "4"]),
    ("t_ycrcb", 64, 64, "natural", 19, "2D-DCT.py", ["-t", "YCrCb"]),
    ("b8_q5_noise", 72, 88, "noise", 20, "2D-DCT.py", ["-q", "5"]),
    ("sa_ycocg", 48, 40, "natural", 21, "YCoCg.py", ["-q", "8"]),
    ("sa_ycrcb", 48, 40, "natural", 22, "YCrCb.py", ["-q", "8"]),
    # -L: the in-process block-size search (src/2D-DCT.py:533-579); the J of every block size is read from the
    # reference's own debug log (:576).  Bright frames and a small step make the indices leave [-128, 127].
    ("L_q4_bright", 128, 128, "natural", 31, "2D-DCT.py", ["-L", "50", "-q", "4", "-g"]),
    ("L_q32", 128, 256, "natural", 32, "2D-DCT.py", ["-L", "2000", "-q", "32", "-g"]),
    # -f: the decoder hands the UN-CLIPPED float64 image to the denoising filter of the class chain (:461) and clips
    # afterwards (:466); src/gaussian_blur.py is cv2.GaussianBlur(y, (5, 5), 0).  Decode-only flag.
    ("f_gaussian_q16", 72, 104, "natural", 33, "2D-DCT.py", ["-q", "16", "-f", "gaussian_blur"]),
]


def run(script, mode, args, env):
    pre = ["-g"] if "-g" in args else []           # -g / --debug belongs to the top-level parser (src/parser.py:75)
    cmd = [sys.executable, script] + pre + [mode] + [x for x in args if x != "-g"]
    r = subprocess.run(cmd, cwd=REF_SRC, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed:\n{r.stdout}\n{r.stderr}")
    return r.stdout + r.stderr


def main():
    if not os.path.isdir(REF_SRC):
        sys.exit("reference not present; golden vectors can only be regenerated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(HERE, "shims") + os.pathsep + env.get("PYTHONPATH", "")
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    for name, H, W, kind, seed, script, flags in CASES:
        tmp = tempfile.mkdtemp(prefix="vcfgold_")
        try:
            img = O.synthetic_frame(H, W, seed, kind)
            if name == "L_q4_bright":
                img = np.clip(img.astype(np.int16) + 70, 0, 255).astype(np.uint8)
            src = os.path.join(tmp, "original.png")
            enc = os.path.join(tmp, "encoded")
            dec = os.path.join(tmp, "decoded.png")
            cv2.imwrite(src, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
            io_flags = ["-c", "z_lib"]
            # Every reference entry point ignores -o/-e/-d on this path and uses
            # its hard-coded defaults (src/2D-DCT.py:374,:470; src/YCoCg.py:35):
            # /tmp/original.png -> /tmp/encoded{.npz,_shape.bin} -> /tmp/decoded.png
            for f in ("/tmp/encoded.npz", "/tmp/encoded_shape.bin", "/tmp/decoded.png"):
                if os.path.exists(f):
                    os.remove(f)
            shutil.copy(src, "/tmp/original.png")
            eflags = [f for i, f in enumerate(flags) if f != "-f" and (i == 0 or flags[i - 1] != "-f")]
            log = run(script, "encode", eflags + io_flags, env)
            dflags = [f for i, f in enumerate(flags) if f not in ("-L", "-g") and (i == 0 or flags[i - 1] != "-L")]
            if "-L" in flags:        # the decoder does not know which block size the encoder chose (:64-66): tell it
                chosen = int(log.split("optimal block_size=")[1].split()[0])
                dflags = dflags + ["-B", str(chosen)]
            run(script, "decode", dflags + io_flags, env)
            shutil.copy("/tmp/encoded.npz", enc + ".npz")
            shutil.copy("/tmp/decoded.png", dec)
            if script == "2D-DCT.py":
                shape_bin = np.frombuffer(open("/tmp/encoded_shape.bin", "rb").read(), dtype=np.int32)
            else:
                shape_bin = np.array([H, W, 3], dtype=np.int32)
            idx = np.load(enc + ".npz")["a"]
            out = cv2.cvtColor(cv2.imread(dec, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)
            extra = {}
            if "-L" in flags:
                Bs, Js = [], []
                for line in log.splitlines():
                    if "J=" in line and "block_size=" in line:
                        Js.append(float(line.split("J=")[1].split()[0]))
                        Bs.append(int(line.split("block_size=")[1].split()[0]))
                extra["L_block_sizes"] = np.array(Bs, dtype=np.int64)
                extra["L_J"] = np.array(Js, dtype=np.float64)
                extra["L_chosen"] = np.int64(chosen)
                extra["L_lambda"] = np.float64(flags[flags.index("-L") + 1])
            if name == "default_96x80":
                txt = subprocess.run(
                    [sys.executable, "RDE.py", "-o", src, "-c", enc + ".npz", "-d", dec],
                    cwd=REF_SRC, env=env, capture_output=True, text=True).stdout
                for line in txt.splitlines():
                    if "Distortion (RMSE)" in line:
                        extra["rde_rmse_2dp"] = np.float64(line.split(":")[-1])
                    if line.startswith("RDE: J"):
                        extra["rde_J_2dp"] = np.float64(line.split("=")[-1])
                    if "Code-stream:" in line:
                        extra["rde_codestream_bytes"] = np.int64(line.split("]")[-1].split()[0])
            np.savez_compressed(
                os.path.join(GOLD, f"ref_flow_{name}.npz"),
                img=img, idx=idx, decoded=out, shape_bin=shape_bin,
                script=np.array(script), flags=np.array(flags, dtype="U16"), **extra)
            print(f"{name}: idx {idx.shape} {idx.dtype}, decoded {out.shape}, extra {extra}")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
