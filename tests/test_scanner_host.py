"""Host-side pieces of the drop-in Scanner that need no GPU: the one-pass check of a provider's ready lists
(scanner._checked_lists) against bpprobs.to_coo window by window."""
import numpy as np
import pytest

from rmdetect_b200.bpprobs import to_coo
from rmdetect_b200.scanner import _Coo, _checked_lists
from rmdetect_b200.synth import SynthFold, synth_sequence


def u16(a):
    return np.array(a, dtype=np.uint16)


def f64(a):
    return np.array(a, dtype=np.float64)


def test_ready_lists_pass_untouched():
    good = [(u16([0, 0, 2]), u16([1, 5, 3]), f64([.1, .2, .3])), (u16([]), u16([]), f64([])), (u16([1]), u16([2]), f64([.5])),
            (u16([0, 1]), u16([3, 2]), f64([.5, .5]))]                  # a window may start below the previous window's last pair
    out = list(_checked_lists(good, [6, 5, 4, 4]))
    assert all(a is b for a, b in zip(out, good))
    lead = [(u16([]), u16([]), f64([])), (u16([0]), u16([1]), f64([.1])), (u16([]), u16([]), f64([]))]
    out = list(_checked_lists(lead, [4, 4, 4]))
    assert len(out) == 3 and out[1] is lead[1]
    assert len(list(_checked_lists([(u16([]), u16([]), f64([]))], [4]))) == 1


@pytest.mark.parametrize("lists,win_lens", [
    ([(u16([0, 0]), u16([5, 1]), f64([.1, .2]))], [6]),                                            # not sorted
    ([(u16([0]), u16([7]), f64([.1]))], [6]),                                                      # outside the window
    ([(u16([0, 1]), u16([2, 3]), f64([0.0, .1]))], [6]),                                           # a stored zero
    ([(u16([0]), u16([1]), f64([.1])), (u16([2, 1]), u16([3, 2]), f64([.1, .2]))], [4, 4]),        # the second window unsorted
    ([([0], [1], [0.5])], [4]),                                                                    # plain Python lists
    ([(u16([3]), u16([1]), f64([.1]))], [6]),                                                      # lower triangle
])
def test_anything_else_goes_through_to_coo(lists, win_lens):
    got = list(_checked_lists(lists, win_lens))
    want = [to_coo(_Coo(i, j, p), n) for (i, j, p), n in zip(lists, win_lens)]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(g, w))


def test_a_repeated_pair_is_still_refused():
    with pytest.raises(ValueError):
        list(_checked_lists([(u16([0, 0]), u16([5, 5]), f64([.1, .2]))], [6]))


def test_the_stand_in_fold_is_in_the_ready_form():
    fold = SynthFold(3)
    seq = synth_sequence(400, 3)
    wins = [seq[s:s + 150] for s in (0, 75, 150, 250)] + [seq[:40]]
    lists = [fold.coo(w) for w in wins]
    out = list(_checked_lists(lists, [len(w) for w in wins]))
    assert all(a is b for a, b in zip(out, lists))
