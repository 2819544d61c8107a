"""TEST INFRASTRUCTURE ONLY -- ``chess.syzygy`` import stub (python-chess, un-vendored).

The reference ships only LFS pointer stubs under ``data/syzygy``, so tablebases are
effectively absent; ``internal.py:257`` catches the failure and plays on without adjudication.
"""


def open_tablebase(path):
    raise FileNotFoundError(f"syzygy tablebases are not supported by the oracle shim: {path}")
