"""Host-only parts of the encoding interface (no GPU): SURVEY 8a row E6, the 73-plane permutations of the training augmentation."""
import json
import os

import numpy as np


def test_flip_and_rotate_permutations_equal_the_reference_arrays(golden_dir):
    """build_horizontal_flip_permutation / build_rotate180_permutation == the arrays of the unmodified reference
    (azchess/encoding.py:310-386; tests/golden/permutations_golden.json from make_permutation_golden.py), and they are involutions
    (the reference's own test, tests/test_encoding_random.py:20-27)."""
    from matrix0_b200 import encoding as enc
    with open(os.path.join(golden_dir, "permutations_golden.json")) as f:
        g = json.load(f)
    hp, rp = enc.build_horizontal_flip_permutation(), enc.build_rotate180_permutation()
    assert np.asarray(hp).dtype == np.int64 and np.asarray(rp).dtype == np.int64
    assert list(map(int, hp)) == g["horizontal_flip"] and list(map(int, rp)) == g["rotate180"]
    assert (hp[hp] == np.arange(73)).all() and (rp[rp] == np.arange(73)).all()
