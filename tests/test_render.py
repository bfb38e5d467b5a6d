"""the UI's colour map (src/app.rs:235-404) on the device, against a numpy f32 restatement."""
import numpy as np
import pytest

from cfd_demo_b200.types import Cylinder, Grid, Scenario, SimulationParams, VelocityScheme, default_grid
from oracle import colormap


def test_colormap_restatement_known_answers():
    """Hand-checkable cases of the numpy restatement: linear pressure ramp -> red/blue ramp with Rust's truncating
    casts, constant field -> `max = min + 1` (:247-249) -> pure blue, cylinder cells grey (:262-268)."""
    g = Grid.uniform(8, 4, 8.0, 4.0, Cylinder(2.5, 1.5, 0.6))
    p = np.tile(np.arange(8, dtype=np.float32), (4, 1))
    u = np.zeros((4, 9), dtype=np.float32)
    v = np.zeros((5, 8), dtype=np.float32)
    img, lo, hi = colormap.render(0, p, u, v, g)
    assert (lo, hi) == (0.0, 7.0)
    assert img[0, 0].tolist() == [0, 0, 255, 255] and img[0, 7].tolist() == [255, 0, 0, 255]
    assert img[3, 3].tolist() == [int(np.float32(3) / np.float32(7) * np.float32(255)), 0,
                                  int((np.float32(1) - np.float32(3) / np.float32(7)) * np.float32(255)), 255]
    assert img[1, 2].tolist() == [128, 128, 128, 255]  # cell centre (2.5, 1.5) is the cylinder's centre
    img, lo, hi = colormap.render(0, np.full((4, 8), 3.0, np.float32), u, v, g)
    assert (lo, hi) == (3.0, 3.0) and img[0, 0].tolist() == [0, 0, 255, 255]
    # rigid rotation u = -y, v = x has vorticity 2 on interior cells, 0 on the ring
    g2 = Grid.uniform(8, 8, 8.0, 8.0, None)
    yy = (np.arange(8, dtype=np.float32) + np.float32(0.5))[:, None]
    xx = (np.arange(8, dtype=np.float32) + np.float32(0.5))[None, :]
    u2 = np.broadcast_to(-yy, (8, 9)).astype(np.float32)
    v2 = np.broadcast_to(xx, (9, 8)).astype(np.float32)
    w = colormap.mapped_quantity(2, np.zeros((8, 8), np.float32), u2, v2, 8, 8, g2.dx, g2.dy)
    assert np.all(w[1:-1, 1:-1] == 2.0) and not w[0].any() and not w[:, 0].any()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_device_colour_map_matches_app_rs_restatement(precision):
    from cfd_demo_b200.model import Model
    g = default_grid()
    m = Model(g, SimulationParams(velocity_scheme=VelocityScheme.SecondOrder), precision=precision)
    for _ in range(30):
        m.update()
    snap = m.get_snapshot()
    for mode in (0, 1, 2):
        ref, lo, hi = colormap.render(mode, snap.p, snap.u, snap.v, g)
        img, glo, ghi = m.render_rgba(mode)
        assert (glo, ghi) == (lo, hi) and hi > lo, (mode, glo, ghi, lo, hi)
        assert img.shape == ref.shape and np.array_equal(img, ref), (mode, int((img != ref).sum()))
        assert (img[..., 0] == 128).any() and img[..., 0].max() == 255  # cylinder overlay and the full ramp


@pytest.mark.gpu
def test_device_colour_map_constant_field_and_pinned_destination():
    from cfd_demo_b200.model import Model, PinnedBuffer
    g = Grid.uniform(64, 32, 1.0, 0.5, None)
    m = Model(g, SimulationParams(scenario=Scenario.Cavity))
    img, lo, hi = m.render_rgba(0)           # all-zero pressure: range 0 -> max = min + 1 -> pure blue
    assert (lo, hi) == (0.0, 0.0) and np.all(img == np.array([0, 0, 255, 255], dtype=np.uint8))
    for _ in range(5):
        m.update()
    buf = PinnedBuffer(64 * 32 * 4, np.uint8)
    a, _, _ = m.render_rgba(1, out=buf)
    b, _, _ = m.render_rgba(1)
    assert np.array_equal(a, b)
    snap = m.get_snapshot()
    ref, _, _ = colormap.render(1, snap.p, snap.u, snap.v, g)
    assert np.array_equal(b, ref)
