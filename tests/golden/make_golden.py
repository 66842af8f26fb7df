"""Generates tests/golden/* from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

Inputs  : /root/reference/testdata/*.png, /root/reference/hardware/mems/*, hardware/exp_disp.mem,
          and the reference's own C golden generator hardware/sad.c compiled UNMODIFIED into
          oracle/_ref/hw_sad (oracle/Makefile).
Outputs : gray fixtures produced by the Go-exact loader (oracle/go_image.py), oracle outputs and
          their SHA-256, the FPGA known-answer vectors, and hw_sad runs on random patches.
The GPU box has no /root/reference: tests read only what this script wrote.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from oracle.go_image import load_png    # noqa: E402

REF = "/root/reference"
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save_gray(name, arr):
    Image.fromarray(arr, "L").save(os.path.join(HERE, name), optimize=True)


def main():
    O.build()
    manifest = {"pairs": {}, "survey_pins": {}}
    # --- cfg1: the four 640x480 RGBA pairs through LoadPNG semantics ---------------------
    for tag in ("00001", "00002", "00335", "01000"):
        L = load_png(f"{REF}/testdata/L_{tag}.png")
        R = load_png(f"{REF}/testdata/R_{tag}.png")
        save_gray(f"L_{tag}_gray.png", L)
        save_gray(f"R_{tag}_gray.png", R)
        out = O.frame_box(L, R, 9, 64)
        save_gray(f"disp_{tag}_b9_d64.png", out)
        manifest["pairs"][tag] = {"left_sha256": sha(L), "right_sha256": sha(R),
                                  "b9_d64_sha256": sha(out)}
    L = load_png(f"{REF}/testdata/L_00001.png"); R = load_png(f"{REF}/testdata/R_00001.png")
    manifest["pairs"]["00001"]["b16_d64_sha256"] = sha(O.frame_box(L, R, 16, 64))
    # spot-check the closed form against the literal restatement on the real pair
    lit = O.frame_literal_mt(L, R, 9, 64, threads=os.cpu_count() or 1, y0=200, y1=216)
    assert np.array_equal(lit, O.frame_box(L, R, 9, 64)[200:216]), "literal != box on cfg1 rows"
    # --- cfg2: 1920x1080 RGB pair; LoadPNG as written gives zeros, keep the intended luma --
    for nm in ("im0", "im1"):
        z = load_png(f"{REF}/testdata/{nm}.png", "loadpng")
        assert not z.any(), "LoadPNG RGBA>>24 bug should yield all-zero"
        save_gray(f"{nm}_intended_gray.png", load_png(f"{REF}/testdata/{nm}.png", "intended"))
    L = load_png(f"{REF}/testdata/im0.png", "intended"); R = load_png(f"{REF}/testdata/im1.png", "intended")
    out = O.frame_box(L, R, 15, 256)
    save_gray("disp_im0_im1_intended_b15_d256.png", out)
    manifest["cfg2"] = {"left_sha256": sha(L), "right_sha256": sha(R), "b15_d256_sha256": sha(out),
                        "loadpng_faithful_output_sha256": sha(np.zeros((1080, 1920), np.uint8))}
    Lb = load_png(f"{REF}/testdata/im0-bs.png", "intended"); Rb = load_png(f"{REF}/testdata/im1-bs.png", "intended")
    manifest["cfg2"]["bs_b15_d256_sha256"] = sha(O.frame_box(Lb, Rb, 15, 256))
    # --- SURVEY.md §8(c) pins, recorded verbatim so the tests can compare -----------------
    manifest["survey_pins"] = {
        "00001_b9_d64": "0bde916590100d42d2fb6c0aca7081ba4f912c2b03a8f3bb5a8d9946454317fb",
        "00001_b16_d64": "f73ec2b08bf5feb0f7d71a2bbb1187f7bcac35bd99d6fc99e82c7c9be503ee0f",
        "00002_b9_d64": "d9b4cd026070c46be2dc5505b8c8b2677d281cb0b7106ef557a0dcad91503289",
        "00335_b9_d64": "500471fa64ff759141e80f41e3e5bf18ed5158a7de52c8d07cc2bb4116195860",
        "01000_b9_d64": "20aa2b998ce7f1de0e73268f30ec7a8befbb1b9f91542044dbb0bc6515bfa6f2",
        "im0_im1_intended_b15_d256": "c2f108c0b3f65208ed35db5a6c826c5749028d23d23cfe6c7f23f30f34bc0006",
        "im0bs_im1bs_intended_b15_d256": "4f94f18c9296ccc98293024b846b9ef56d71bc1d14dcaf6690872ec26e26fb5e",
        "zeros_1920x1080": "11283ef755895422e6f28b93f3d78cad7539891cf2893c9fdccefb923c5bf70b",
    }
    # --- FPGA known answers: hardware/mems patches, exp_disp_p.mem (test.py), exp_disp.mem (sad.c)
    Lp = np.stack([np.fromfile(f"{REF}/hardware/mems/img_L_patch_{k}.raw", np.uint8).reshape(128, 128) for k in range(4)])
    Rp = np.stack([np.fromfile(f"{REF}/hardware/mems/img_R_patch_{k}.raw", np.uint8).reshape(128, 128) for k in range(4)])
    rd = lambda p: np.array([int(t, 16) for t in open(p).read().split()], np.uint8)
    exp_p = rd(f"{REF}/hardware/mems/exp_disp_p.mem").reshape(4, 128, 128)
    exp_c = rd(f"{REF}/hardware/exp_disp.mem").reshape(128, 128)
    # --- the reference's C generator, run here on random patches ---------------------------
    hw = os.path.join(ROOT, "oracle", "_ref", "hw_sad")
    rnd_L, rnd_R, rnd_out = [], [], []
    rng = np.random.default_rng(20261018)
    for k in range(3):
        base = rng.integers(0, 256, (128, 128 + 128), dtype=np.uint8)
        shift = 5 + 7 * k
        l = base[:, 64:192].copy()
        r = base[:, 64 + shift:64 + shift + 128].copy() if k else rng.integers(0, 256, (128, 128), dtype=np.uint8)
        if k == 2:
            l = (l // 64) * 64   # tie-heavy
            r = (r // 64) * 64
        with tempfile.TemporaryDirectory() as td:
            os.mkdir(os.path.join(td, "mems"))
            l.tofile(os.path.join(td, "mems", "img_L_patch_1.raw"))
            r.tofile(os.path.join(td, "mems", "img_R_patch_1.raw"))
            subprocess.check_call([hw], cwd=td)
            rnd_out.append(rd(os.path.join(td, "exp_disp.mem")).reshape(128, 128))
        rnd_L.append(l); rnd_R.append(r)
    np.savez_compressed(os.path.join(HERE, "fpga_vectors.npz"), L=Lp, R=Rp, exp_disp_p=exp_p, exp_disp_c_patch1=exp_c,
                        hw_sad_L=np.stack(rnd_L), hw_sad_R=np.stack(rnd_R), hw_sad_out=np.stack(rnd_out))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    for k, v in manifest["survey_pins"].items():
        print(k, v[:12])
    print("written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
