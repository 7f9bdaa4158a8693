"""Seeded VarDCT codestreams for the parity tests, written by the oracle's plain encoder (oracle/jxlo_encode.h):
the reference ships no VarDCT fixture larger than 40x50 and libjxl cannot be built here."""
import functools

import numpy as np

import jxlo
from conftest import read_golden


@functools.lru_cache(maxsize=1)
def natural_image():
    """The reference's bench image (2122x1433), decoded by the oracle from the golden bench.jxl."""
    return jxlo.decode(read_golden("bench.jxl"), 3, jxlo.UINT8)


def crop(h, w, y0=0, x0=0):
    return np.ascontiguousarray(natural_image()[y0:y0 + h, x0:x0 + w])


def frame_4k():
    """3840x2160 RGB8 with natural-image statistics: the bench image tiled and cropped."""
    img = natural_image()
    reps = (2160 + img.shape[0] - 1) // img.shape[0], (3840 + img.shape[1] - 1) // img.shape[1]
    return np.ascontiguousarray(np.tile(img, (reps[0], reps[1], 1))[:2160, :3840])


def synthetic(h, w, seed):
    """Smooth gradients + noise + hard edges (procedural, seeded)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(x / 37.0 + seed) * np.cos(y / 23.0), 128 + 90 * np.cos((x + y) / 51.0),
                    128 + 80 * np.sin(y / 17.0)], axis=2)
    img += rng.normal(0, 6, img.shape)
    for _ in range(12):
        x0, y0 = rng.integers(0, w), rng.integers(0, h)
        img[y0:y0 + rng.integers(4, 60), x0:x0 + rng.integers(4, 60)] = rng.integers(0, 255, 3)
    return np.clip(img, 0, 255).astype(np.uint8)


# (name, image factory, encoder arguments)
SMALL_CASES = [
    ("dct8_plain", lambda: crop(300, 400), dict(strategy_mode=0, gab=False, epf_iters=0, dc_smoothing=False)),
    ("dct8_filters", lambda: crop(300, 400, 100, 200), dict(strategy_mode=0, gab=True, epf_iters=2)),
    ("heuristic", lambda: crop(520, 700, 300, 500), dict(strategy_mode=2)),
    ("all_strategies", lambda: crop(300, 520, 600, 900), dict(strategy_mode=1, random_side_info=True, epf_iters=3, seed=3)),
    ("three_passes", lambda: crop(333, 517, 50, 1000), dict(strategy_mode=1, random_side_info=True, epf_iters=1, num_passes=3, seed=7)),
    ("synthetic_d3", lambda: synthetic(280, 264, 5), dict(strategy_mode=2, distance=3.0)),
    ("synthetic_d05", lambda: synthetic(264, 300, 9), dict(strategy_mode=2, distance=0.5, gab=False, epf_iters=1)),
    ("odd_size", lambda: crop(257, 263, 700, 100), dict(strategy_mode=1, seed=11)),
    # prefix codes instead of ANS in every stream of the frame (tree, DC / metadata, coefficient orders, AC passes)
    ("prefix_codes", lambda: crop(300, 400, 100, 200), dict(strategy_mode=1, random_side_info=True, seed=3, epf_iters=1, prefix_codes=True)),
    ("prefix_codes_two_passes", lambda: crop(280, 330, 500, 300), dict(strategy_mode=2, num_passes=2, prefix_codes=True)),
    # frame upsampling 2x / 4x / 8x (stage_upsampling.cc): the encoder is handed the low-resolution frame, the image is
    # `upsampling` times its size; sizes that are not multiples of 8 (mirrored edges, kernels at the last column)
    ("upsampling_2", lambda: crop(150, 203, 100, 200), dict(strategy_mode=2, upsampling=2)),
    ("upsampling_4", lambda: crop(67, 90, 300, 800), dict(strategy_mode=1, random_side_info=True, seed=4, epf_iters=1, upsampling=4)),
    ("upsampling_8", lambda: crop(33, 41, 640, 960), dict(strategy_mode=2, gab=False, epf_iters=0, upsampling=8)),
]
STRATEGY_CASES = [("strategy_%d" % s, lambda: crop(264, 520, 400, 300),
                   dict(strategy_mode=100 + s, gab=False, epf_iters=0, dc_smoothing=False)) for s in range(27)]


@functools.lru_cache(maxsize=None)
def encoded(name):
    for n, make, kw in SMALL_CASES + STRATEGY_CASES:
        if n == name:
            img = make()
            u = kw.get("upsampling", 1)
            return jxlo.encode_vardct(img, **kw), (img.shape[0] * u, img.shape[1] * u)
    raise KeyError(name)
