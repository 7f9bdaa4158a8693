"""Seeded lossless Modular streams written by the oracle's plain encoder (oracle/jxlo_enc_modular.h): the decoder
branches the reference's own fixtures do not reach -- palette, delta palette (explicit and implicit entries), squeeze,
every predictor and property through fixed and random trees, prefix codes and LZ77 -- shared by the CPU emulation tests
and the GPU parity tests."""
import functools

import numpy as np

import jxlo


def smooth(h, w, c, bits, seed=1):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    chans = []
    for k in range(c):
        v = (np.sin(x / (7.0 + k)) * 0.3 + np.cos(y / (11.0 + 2 * k)) * 0.3 + 0.5) * ((1 << bits) - 1)
        v = v + rng.integers(-3, 4, size=v.shape)
        chans.append(np.clip(v, 0, (1 << bits) - 1))
    return np.stack(chans, -1).astype(np.uint16)


def few_colors(h, w, c=3, seed=2):
    """At most six colours, with long runs (a flat left half): palette + LZ77 material."""
    rng = np.random.default_rng(seed)
    base = np.array([40, 30, 20, 50])[:c]
    img = (rng.integers(0, 6, size=(h, w, 1)) * base[None, None, :]).astype(np.uint16)
    img[:, :w // 2] = img[:1, :w // 2]
    return img


def ramps(h, w, seed=3):
    """Smooth ramps of few colours: a delta palette reproduces most pixels as prediction + delta."""
    y, x = np.mgrid[0:h, 0:w]
    r = (x // 4) % 32 * 8
    g = (y // 4) % 32 * 8
    b = ((x + y) // 8) % 16 * 16
    return np.stack([r, g, b], -1).astype(np.uint16)


# name -> (image factory, encoder keywords, (num_channels, data type) of the decode)
CASES = {
    "one_leaf_gradient": (lambda: smooth(50, 70, 3, 8), dict(), (3, jxlo.UINT8)),
    "gradient_tree": (lambda: smooth(50, 70, 3, 8), dict(tree=1), (3, jxlo.UINT8)),
    "wp_tree": (lambda: smooth(50, 70, 3, 8), dict(tree=2), (3, jxlo.UINT8)),
    "random_tree_a": (lambda: smooth(61, 83, 3, 8, 4), dict(tree=3, seed=5), (3, jxlo.UINT8)),
    "random_tree_b": (lambda: smooth(77, 45, 3, 8, 5), dict(tree=3, seed=11, rct=6), (4, jxlo.UINT8)),
    "random_tree_c": (lambda: smooth(64, 64, 4, 16, 6), dict(bits=16, alpha=True, tree=3, seed=9, rct=6), (4, jxlo.UINT16)),
    "rct_17": (lambda: smooth(50, 70, 3, 8), dict(rct=17, tree=1), (3, jxlo.UINT8)),
    "rct_41": (lambda: smooth(33, 47, 3, 8, 7), dict(rct=41), (3, jxlo.UINT8)),
    "prefix": (lambda: smooth(50, 70, 3, 8), dict(prefix=True), (3, jxlo.UINT8)),
    "prefix_tree": (lambda: smooth(90, 100, 3, 8, 8), dict(prefix=True, tree=1), (3, jxlo.UINT8)),
    "lz77": (lambda: few_colors(60, 80), dict(lz77=True), (3, jxlo.UINT8)),
    "lz77_prefix": (lambda: few_colors(60, 80, seed=9), dict(lz77=True, prefix=True, tree=1), (3, jxlo.UINT8)),
    "lz77_min_symbol_512": (lambda: few_colors(40, 90, seed=10), dict(lz77=True, prefix=True, lz77_min_symbol=512), (3, jxlo.UINT8)),
    "squeeze": (lambda: smooth(50, 70, 3, 8), dict(squeeze=True), (3, jxlo.UINT8)),
    "squeeze_wp_prefix": (lambda: smooth(53, 71, 3, 8, 12), dict(squeeze=True, tree=2, prefix=True), (3, jxlo.UINT8)),
    "squeeze_grey16": (lambda: smooth(90, 100, 1, 16, 13), dict(bits=16, squeeze=True, tree=1), (1, jxlo.UINT16)),
    "palette": (lambda: few_colors(60, 80), dict(palette_colors=16), (3, jxlo.UINT8)),
    "palette_rgba": (lambda: few_colors(40, 50, 4, 14), dict(palette_colors=16, alpha=True), (4, jxlo.UINT8)),
    "palette_grey": (lambda: few_colors(40, 50, 1, 15), dict(palette_colors=8), (1, jxlo.UINT8)),
    "delta_palette_gradient": (lambda: ramps(64, 96), dict(palette_colors=1024, palette_deltas=3, palette_predictor=5), (3, jxlo.UINT8)),
    "delta_palette_weighted": (lambda: ramps(48, 80), dict(palette_colors=1024, palette_deltas=2, palette_predictor=6, lz77=True), (3, jxlo.UINT8)),
    "delta_palette_implicit": (lambda: ramps(40, 64), dict(palette_colors=1024, palette_predictor=4), (3, jxlo.UINT8)),
    "groups": (lambda: smooth(300, 520, 3, 8, 16), dict(), (3, jxlo.UINT8)),
    "groups_random_tree": (lambda: smooth(300, 520, 3, 8, 17), dict(tree=3, seed=2), (3, jxlo.UINT8)),
    "groups_squeeze": (lambda: smooth(300, 520, 3, 8, 18), dict(squeeze=True), (3, jxlo.UINT8)),
    "groups_squeeze_lz77_128": (lambda: smooth(260, 300, 3, 8, 19), dict(squeeze=True, lz77=True, group_size_shift=0), (3, jxlo.UINT8)),
    "groups_rct_prefix_128": (lambda: smooth(200, 300, 3, 8, 20), dict(rct=10, prefix=True, group_size_shift=0), (3, jxlo.UINT8)),
    "groups_palette": (lambda: few_colors(280, 300, 3, 21), dict(palette_colors=16, lz77=True), (3, jxlo.UINT8)),
    "grey_alpha16_wp": (lambda: smooth(90, 100, 2, 16, 22), dict(bits=16, alpha=True, tree=2), (2, jxlo.UINT16)),
}


@functools.lru_cache(maxsize=None)
def encoded(name):
    make, kw, _ = CASES[name]
    img = make()
    return jxlo.encode_modular(img, **kw), img
