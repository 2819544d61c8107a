"""Host logic of `_native.on_own_device` (no GPU needed): which methods are wrapped and that they run inside
`torch.cuda.device(self.device)`; the GPU side is tests/test_selfplay_gpu.py::test_search_on_a_second_device_of_the_same_process."""
import contextlib

import torch

from matrix0_b200 import _native


def test_on_own_device_wraps_plain_methods_only(monkeypatch):
    entered = []

    @contextlib.contextmanager
    def fake_device(dev):
        entered.append(("enter", dev))
        try:
            yield
        finally:
            entered.append(("exit", dev))

    monkeypatch.setattr(torch.cuda, "device", fake_device)

    @_native.on_own_device
    class Thing:
        def __init__(self):
            self.touch()                    # before the device is known: runs unscoped
            self.device = "cuda:1"

        def touch(self):
            return "touched"

        def fails(self):
            raise ValueError("boom")

        def gen(self):
            yield 1

        @property
        def prop(self):
            return 7

        @staticmethod
        def static():
            return 3

        def close(self):
            return "closed"

    t = Thing()
    assert entered == []
    assert t.touch() == "touched" and entered == [("enter", "cuda:1"), ("exit", "cuda:1")]
    assert Thing.touch.__name__ == "touch"
    entered.clear()
    try:
        t.fails()
    except ValueError:
        pass
    assert entered == [("enter", "cuda:1"), ("exit", "cuda:1")]     # the scope is left on errors too
    entered.clear()
    assert list(t.gen()) == [1] and t.prop == 7 and Thing.static() == 3 and t.close() == "closed"
    assert entered == []                                             # generators, properties, static methods, close: untouched


def test_engine_classes_are_scoped():
    from matrix0_b200.arena import _TwoEvaluatorGames
    from matrix0_b200.engine import SearchEngine
    from matrix0_b200.selfplay import SelfPlayEngine
    for cls, names in ((SearchEngine, ("begin", "expand_backup", "set_boards", "status", "select_multi")),
                       (SelfPlayEngine, ("begin_move", "search_step", "end_move", "play_move", "finished_games", "_forward")),
                       (_TwoEvaluatorGames, ("begin_move", "_forward"))):
        for n in names:
            assert hasattr(vars(cls)[n], "__wrapped__"), (cls.__name__, n)
    assert not hasattr(vars(SearchEngine)["__init__"], "__wrapped__")
    assert isinstance(vars(SearchEngine)["bytes"], property)
