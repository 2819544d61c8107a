"""TEST INFRASTRUCTURE ONLY -- ``chess.polyglot`` import stub (python-chess, un-vendored).

``azchess/selfplay/internal.py:71-88`` only opens a book when a path is configured; none ships.
"""


def open_reader(path):
    raise FileNotFoundError(f"polyglot books are not supported by the oracle shim: {path}")
