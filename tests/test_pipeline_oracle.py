"""CPU checks of the input-pipeline oracle (oracle/pipeline_oracle.py): the nearest-neighbour index
against its exact rational definition, structural properties of the composition, and the committed
golden fixture."""
import os

import numpy as np

from oracle import pipeline_oracle as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pipeline_small.npz")


def test_nearest_index_against_exact_rational_definition():
    """TF evaluates floor((o + 1/2) * in / out) in float32.  Against the exact rational value
    ((2o+1)*in) // (2*out) the float32 form is identical except where the exact quotient is an integer
    (a source-pixel boundary), where rounding of in/out may land one below; it is never off elsewhere
    and never by more than one.  The oracle and the kernel both follow the float32 form."""
    n_boundary = 0
    for out in (256, 286, 512, 542, 16, 46):
        o = np.arange(out, dtype=np.int64)
        for src in list(range(1, 1201)) + [1920, 2048, 4096]:
            num = (2 * o + 1) * src
            exact = np.minimum(num // (2 * out), src - 1)
            got = P.nearest_index(out, src)
            diff = np.nonzero(got != exact)[0]
            n_boundary += len(diff)
            assert np.all(exact[diff] - got[diff] == 1), (out, src)
            assert np.all(num[diff] % (2 * out) == 0), (out, src)
    assert n_boundary < 200          # 70 (size pair, index) cases in ~7,200 size pairs


def test_nearest_index_properties():
    for out, src in [(286, 256), (256, 300), (542, 512), (256, 256)]:
        idx = P.nearest_index(out, src)
        assert idx[0] == 0 or src > out
        assert idx.min() >= 0 and idx.max() <= src - 1
        assert np.all(np.diff(idx) >= 0)
    assert np.array_equal(P.nearest_index(256, 256), np.arange(256))       # same size: identity


def test_normalize_endpoints_and_range():
    x = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)
    y = P.normalize(x)
    assert y.dtype == np.float32 and y.min() == -1.0 and y.max() == 1.0
    assert y[0, 0, 0] == np.float32(-1.0) and y[15, 15, 0] == np.float32(1.0)


def test_composition_properties():
    rng = np.random.default_rng(0)
    pair = rng.integers(0, 256, size=(37, 90, 3), dtype=np.uint8)
    a, b = P.pix2pix_process_train(pair, 'left', 16, 3, 7, False)
    af, bf = P.pix2pix_process_train(pair, 'left', 16, 3, 7, True)
    assert np.array_equal(a[:, ::-1], af) and np.array_equal(b[:, ::-1], bf)           # mirror
    ra, rb = P.pix2pix_process_train(pair, 'right', 16, 3, 7, False)
    assert np.array_equal(ra, b) and np.array_equal(rb, a)                             # orientation swaps the halves
    full = P.normalize(P.resize(P.split_img(P.load(pair))[0], 46, 46))
    assert np.array_equal(a, full[3:19, 7:23])                                          # crop window of the +30 image
    # CycleGAN: the pre-resize to img_size makes the prediction path a plain resize
    im = rng.integers(0, 256, size=(50, 41, 3), dtype=np.uint8)
    assert np.array_equal(P.cyclegan_process_pred(im, 16), P.normalize(P.resize(P.load(im), 16, 16)))


def test_golden_fixture():
    g = np.load(GOLDEN)
    pair, im = g["pair"], g["image"]
    a, b = P.pix2pix_process_train(pair, 'left', 16, int(g["cy"]), int(g["cx"]), bool(g["flip"]))
    assert np.array_equal(a, g["p2p_train_a"]) and np.array_equal(b, g["p2p_train_b"])
    a, b = P.pix2pix_process_pred(pair, 'right', 16)
    assert np.array_equal(a, g["p2p_pred_a"]) and np.array_equal(b, g["p2p_pred_b"])
    assert np.array_equal(P.cyclegan_process_train(im, 16, int(g["cy"]), int(g["cx"]), bool(g["flip"])), g["cyc_train"])
    assert np.array_equal(P.cyclegan_process_pred(im, 16), g["cyc_pred"])
