"""Adds the colour fixture of sadgpu_compute_nrgba to tests/golden/ (run in the build container only):
a 320x160 crop of an 8-bit RGBA testdata pair, chosen so that it contains pixels with alpha < 255 (Go decodes such a PNG to
*image.NRGBA; color.GrayModel.Convert then applies the alpha), the Go-exact luma of both crops and the oracle's disparity map.

    python tests/golden/make_rgba_crop.py
"""
import hashlib, json, os, sys
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O                  # noqa: E402
from oracle.go_image import load_png            # noqa: E402

REF = "/root/reference/testdata"
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
CW, CH = 320, 160


def main():
    O.build()
    best = None
    for tag in ("00001", "00002", "00335", "01000"):
        L = np.array(Image.open(f"{REF}/L_{tag}.png")); R = np.array(Image.open(f"{REF}/R_{tag}.png"))
        assert L.shape[2] == 4 and R.shape[2] == 4
        for y0 in range(0, L.shape[0] - CH + 1, 40):
            for x0 in range(0, L.shape[1] - CW + 1, 40):
                n = int((L[y0:y0 + CH, x0:x0 + CW, 3] < 255).sum() + (R[y0:y0 + CH, x0:x0 + CW, 3] < 255).sum())
                if best is None or n > best[0]:
                    best = (n, tag, x0, y0)
    n, tag, x0, y0 = best
    out = {}
    for side in "LR":
        im = Image.open(f"{REF}/{side}_{tag}.png").crop((x0, y0, x0 + CW, y0 + CH))
        im.save(os.path.join(HERE, f"{side}_{tag}_rgba_crop.png"), optimize=True)
        out[side] = load_png(os.path.join(HERE, f"{side}_{tag}_rgba_crop.png"), "intended")
        full = load_png(f"{REF}/{side}_{tag}.png", "intended")
        assert np.array_equal(out[side], full[y0:y0 + CH, x0:x0 + CW])          # cropping commutes with the per-pixel luma
    disp = O.frame_box(out["L"], out["R"], 9, 64)
    Image.fromarray(disp, "L").save(os.path.join(HERE, f"disp_{tag}_rgba_crop_b9_d64.png"), optimize=True)
    mpath = os.path.join(HERE, "manifest.json")
    m = json.load(open(mpath))
    m["rgba_crop"] = {"tag": tag, "x0": x0, "y0": y0, "w": CW, "h": CH, "pixels_with_alpha_below_255": n,
                      "left_gray_sha256": sha(out["L"]), "right_gray_sha256": sha(out["R"]), "b9_d64_sha256": sha(disp)}
    json.dump(m, open(mpath, "w"), indent=1, sort_keys=True)
    print(m["rgba_crop"])


if __name__ == "__main__":
    main()
