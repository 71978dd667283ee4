"""crop_objects reads pixels the way skimage.io.imread (PIL plugin) hands them to the reference: palette images
expanded to RGB / RGBA, everything else in its stored dtype and channel count. CPU only."""
import numpy as np
from PIL import Image

from miso.object_detection.crop import _imread


def test_imread_modes(tmp_path):
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 256, (20, 30), dtype=np.uint8)
    pal = [int(v) for v in rng.integers(0, 256, 768)]
    p = Image.fromarray(idx, mode="P"); p.putpalette(pal); p.save(tmp_path / "p.png")
    a = _imread(str(tmp_path / "p.png"))
    assert a.shape == (20, 30, 3) and a.dtype == np.uint8
    assert np.array_equal(a, np.asarray(pal, np.uint8).reshape(256, 3)[idx])
    p.save(tmp_path / "pt.png", transparency=0)
    assert _imread(str(tmp_path / "pt.png")).shape == (20, 30, 4)
    g16 = rng.integers(0, 65536, (9, 11), dtype=np.uint16)
    Image.fromarray(g16).save(tmp_path / "g.tif")
    b = _imread(str(tmp_path / "g.tif"))
    assert b.dtype == np.uint16 and np.array_equal(b, g16)
    rgb = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(tmp_path / "c.png")
    c = _imread(str(tmp_path / "c.png"))
    assert np.array_equal(c, rgb) and c.flags.writeable
