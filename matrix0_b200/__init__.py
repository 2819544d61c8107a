"""matrix0_b200 -- B200-native self-play search engine behind Matrix0's Python API.

Drop-in surface (mirrors the reference modules named in SURVEY.md section 8b):
  matrix0_b200.encoding  <->  azchess/encoding.py   (encode_board, move_to_index, MoveEncoder ...)
  matrix0_b200.mcts      <->  azchess/mcts.py       (MCTS, MCTSConfig)
  matrix0_b200.model     <->  azchess/model/resnet.py inference forward (PolicyValueNet evaluator)
  matrix0_b200.inference <->  azchess/selfplay/inference.py (shared-memory evaluation server + client)
  matrix0_b200.selfplay  <->  azchess/selfplay/      (selfplay_worker, batched device self-play)

All compute runs in hand-written sm_100a CUDA kernels reached through the C ABI declared in
include/matrix0_b200.h; nothing here falls back to CPU or stock PyTorch ops.
"""
from ._native import NativeLibraryError, load_library  # noqa: F401

__version__ = "0.1.0"
