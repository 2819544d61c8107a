import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the oracle modules (and matrix0_b200's `import chess` for Move objects) need a `chess` module:
# real python-chess if installed, otherwise the oracle restatement
from oracle import chess_shim  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def random_playout_boards(n_games, max_plies, seed):
    """Boards along seeded random playouts from the start position (oracle shim)."""
    import random
    import chess
    rng = random.Random(seed)
    out = []
    for _ in range(n_games):
        b = chess.Board()
        for _ in range(rng.randint(1, max_plies)):
            if b.is_game_over():
                break
            out.append(b.copy())
            b.push(rng.choice(list(b.legal_moves)))
    return out
