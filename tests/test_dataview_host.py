"""CPU: numpy_dataview host logic, following the reference's test/test_dataview.py
(iteration equals numpy rows :31-46, sub-array fields :49-60, masked :63-76, pickle :79-122,
digest :131-150) and the record layout of runtime_type.hpp:123-134."""
import hashlib
import pickle

import numpy as np
import numpy.ma as ma
import pytest

import common_b200 as cb
from common_b200 import _lib
from common_b200.dataview import get_c_types


def test_iteration_equals_numpy_rows():
    Y = np.array([(True, 2, 3.5), (False, 7, -1.0), (True, 0, 9.25)], dtype=[("", bool), ("", np.int32), ("", np.float32)])
    view = cb.numpy_dataview(Y)
    assert view.size() == 3 and len(view) == 3
    for a, b in zip(view, Y):
        assert a == b
    for a, b in zip(view, Y):  # a second pass restarts from the top
        assert a == b


def test_subarray_fields_become_vectors():
    Y = np.zeros(4, dtype=[("a", bool), ("v", np.float64, (3,)), ("c", np.uint16)])
    Y["v"] = np.arange(12).reshape(4, 3)
    t = cb.numpy_dataview(Y).types()
    assert [(x.prim, x.n, x.vec) for x in t] == [(_lib.TYPE_B, 1, 0), (_lib.TYPE_F64, 3, 1), (_lib.TYPE_U16, 1, 0)]
    data, mask = cb.numpy_dataview(Y).raw()
    assert data.shape == (4, 1 + 24 + 2) and mask is None  # offsets = running sum of sizes
    assert np.frombuffer(data[2, 1:25].tobytes(), np.float64).tolist() == [6.0, 7.0, 8.0]


def test_masked_rows_and_mask_layout():
    Y = ma.array(np.array([(True, 1.5), (False, 2.5)], dtype=[("a", bool), ("b", np.float32)]),
                 mask=[(False, True), (False, False)])
    view = cb.numpy_dataview(Y)
    rows = list(view)
    assert rows[0]["a"] == True and rows[0].mask["b"] and not rows[0].mask["a"]
    assert not hasattr(rows[1], "mask") or not np.any(rows[1].mask.tolist())
    data, mask = view.raw()
    assert mask.tolist() == [[0, 1], [0, 0]]  # one bool per element, running sum of n


def test_pickle_and_digest():
    Y = np.array([(1, 2.0), (3, 4.0)], dtype=[("", np.int64), ("", np.float64)])
    v = cb.numpy_dataview(Y)
    w = pickle.loads(pickle.dumps(v))
    assert [tuple(r) for r in v] == [tuple(r) for r in w]
    assert v.digest().hexdigest() == w.digest().hexdigest()
    Y2 = Y.copy(); Y2[1][1] = 5.0
    assert cb.numpy_dataview(Y2).digest().hexdigest() != v.digest().hexdigest()
    with pytest.raises(NotImplementedError):
        cb.numpy_dataview(ma.array(Y, mask=[(False, True), (False, False)])).digest()


def test_rejects_non_structured():
    with pytest.raises(ValueError):
        cb.numpy_dataview(np.zeros(5))
    with pytest.raises(ValueError):
        cb.numpy_dataview(np.zeros((2, 2), dtype=[("a", bool)]))


def test_permutation_is_a_permutation():
    Y = np.array([(i,) for i in range(20)], dtype=[("", np.int32)])
    v = cb.numpy_dataview(Y)
    v.permute(np.random.default_rng(4382))
    got = [int(r[0]) for r in v]
    assert sorted(got) == list(range(20)) and got != list(range(20))
    v.reset_permutation()
    assert [int(r[0]) for r in v] == list(range(20))


def test_models_mirror_reference_descriptors():
    # test/test_models.py:17-51
    assert cb.nich() is cb.nich and cb.bb() is cb.bb
    assert cb.niw(3).get_np_dtype().shape == (3,)
    for m in (cb.bb, cb.gp, cb.nich, cb.dd(5), cb.niw(4)):
        m2 = pickle.loads(pickle.dumps(m))
        assert m2.name() == m.name() and m2._param() == m._param()
    assert cb.bb.default_hyperparams() == {"alpha": 1.0, "beta": 1.0}
    assert cb.nich.default_hyperparams() == {"mu": 0.0, "kappa": 1.0, "sigmasq": 1.0, "nu": 1.0}
    assert cb.dd(3).default_hyperparams() == {"alphas": [1.0, 1.0, 1.0]}
    with pytest.raises(ValueError):
        cb.dd(0)
