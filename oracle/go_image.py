"""Go-exact PNG -> 8-bit gray loader — TEST INFRASTRUCTURE ONLY.

Restates what the reference does to a PNG before the SAD path sees it:
  * despair.LoadPNG            pkg/despair/png.go:10-40 (type switch :30-37)
  * convertGrayToGray          pkg/despair/gray.go:15-17
  * convertRGBAToGray          pkg/despair/gray.go:20-40  (8-bit channels, >>24  => always 0)
  * convertGenericToGray       pkg/despair/gray.go:43-58  (16-bit channels from color.RGBA())
  * color.GrayModel.Convert    used by OutputCamera, pkg/camera/output.go:145,160 ("intended")
Go stdlib behaviour relied on (image/png, image/color of Go 1.24, not under /root/reference):
  8-bit gray PNG -> *image.Gray; 8-bit RGB (no tRNS) -> *image.RGBA (A=255);
  8-bit RGBA -> *image.NRGBA whose RGBA() is c16 = (c8*0x101)*a8/0xff, a16 = a8*0x101.
PIL is used only to get the raw stored samples; no PIL colour conversion is applied.
"""
import numpy as np
from PIL import Image


def _luma16(r, g, b):
    # gray.go:53-55 / Go color.grayModel: (19595 r + 38470 g + 7471 b + 1<<15) >> 24 on 16-bit channels
    return ((19595 * r + 38470 * g + 7471 * b + (1 << 15)) >> 24).astype(np.uint8)


def load_png(path: str, mode: str = "loadpng") -> np.ndarray:
    """mode='loadpng'  : exactly despair.LoadPNG (including the RGBA >>24 bug);
       mode='intended' : color.GrayModel.Convert of every pixel (what OutputCamera does)."""
    im = Image.open(path)
    if im.mode == "L":
        return np.array(im, np.uint8)
    if im.mode == "RGB":
        a = np.array(im).astype(np.uint64)
        r, g, b = a[..., 0], a[..., 1], a[..., 2]
        if mode == "loadpng":   # *image.RGBA -> convertRGBAToGray on 8-bit values
            return ((19595 * r + 38470 * g + 7471 * b + (1 << 15)) >> 24).astype(np.uint8)
        return _luma16(r * 257, g * 257, b * 257)
    if im.mode == "RGBA":       # *image.NRGBA -> generic path in both modes
        a = np.array(im).astype(np.uint64)
        al = a[..., 3]
        c16 = lambda c: (c * 257 * al) // 255
        return _luma16(c16(a[..., 0]), c16(a[..., 1]), c16(a[..., 2]))
    if im.mode == "LA":         # gray+alpha decodes to *image.NRGBA
        a = np.array(im).astype(np.uint64)
        y16 = (a[..., 0] * 257 * a[..., 1]) // 255
        return _luma16(y16, y16, y16)
    if im.mode in ("I;16", "I;16B", "I"):   # *image.Gray16 -> generic path
        y16 = np.array(im).astype(np.uint64) & 0xFFFF
        return _luma16(y16, y16, y16)
    raise NotImplementedError(f"PNG mode {im.mode!r} is not covered by the Go-exact loader")
