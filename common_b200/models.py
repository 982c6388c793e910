"""Model descriptors -- host mirror of microscopes/models.pyx:96-290.

Same names (bb, gp, nich, dd(size), niw(dim)), same default hyperparameters
(models.pyx:189,211,223,238,264-269), same ``nich() is nich`` idiom
(models.pyx:142-146, test/test_models.py:17-19).  ``c_desc()`` returns the
msb_model_desc the C ABI consumes, in place of the reference's Cython handle
owning a shared_ptr<model> (microscopes/_models.pyx:16-52).
"""
import numpy as np

from . import _lib

_FAMILY = {"dm": _lib.FAMILY_DM, "bbnc": _lib.FAMILY_BBNC, "bb": _lib.FAMILY_BB, "bnb": _lib.FAMILY_BNB, "gp": _lib.FAMILY_GP, "nich": _lib.FAMILY_NICH,
           "dd": _lib.FAMILY_DD, "niw": _lib.FAMILY_NIW}


def _validate_positive(v, param_name):
    # microscopes/common/validator.py: validate_positive
    if v <= 0:
        raise ValueError("need positive value for param `%s'" % param_name)


class model_descriptor(object):
    def __init__(self, name, dtype, default_hyperparams, param=None):
        self._name = name
        self._dtype = np.dtype(dtype)
        self._default_hyperparams = default_hyperparams
        self._param_value = param

    def name(self):
        return self._name

    def get_np_dtype(self):
        """dtype of one value (py_model.get_np_dtype, models.pyx:62-63)"""
        return self._dtype

    def py_desc(self):
        return self

    def c_desc(self):
        return _lib.ModelDesc(_FAMILY[self._name], int(self._param_value or 0))

    def default_hyperparams(self):
        return self._default_hyperparams

    def _param(self):
        return self._param_value

    def hp_keys(self):
        return list(self._default_hyperparams.keys())

    def __reduce__(self):
        return (_reconstruct_model_descriptor, (self._name, self._param()))

    def __call__(self):
        """Make models callable so nich() == nich (models.pyx:142-146)."""
        return self


def _reconstruct_model_descriptor(name, param):
    desc = globals()[name]
    return desc if param is None else desc(param)


# Value types as seen through get_runtime_type (distributions.hpp:399-403): bb bool,
# gp uint32, nich float32, dd int32 [R: SURVEY.md section 2a]
bb = model_descriptor("bb", np.bool_, {"alpha": 1., "beta": 1.})
bbnc = model_descriptor("bbnc", np.bool_, {"alpha": 1., "beta": 1.})            # models.pyx:248-253, src/models/bbnc.cpp
bnb = model_descriptor("bnb", np.uint32, {"alpha": 1., "beta": 1., "r": 1.})   # models.pyx:196-205
gp = model_descriptor("gp", np.uint32, {"alpha": 1., "inv_beta": 1.})
nich = model_descriptor("nich", np.float32, {"mu": 0., "kappa": 1., "sigmasq": 1., "nu": 1.})


def dd(size):
    _validate_positive(size, "size")
    return model_descriptor("dd", np.int32, {"alphas": [1.] * size}, param=int(size))


def dm(categories):
    """Dirichlet-multinomial over count vectors (models.pyx:279-292, src/models/dm.cpp): one value = categories counts"""
    _validate_positive(categories, "categories")
    # value type: dm.hpp:186-190
    return model_descriptor("dm", np.dtype((np.int32, (categories,))), {"alphas": [1.] * categories}, param=int(categories))


def niw(dim):
    _validate_positive(dim, "dim")
    # the reference hands NIW float64 rows (models.pyx:259) that are cast per element
    return model_descriptor(
        "niw", np.dtype((float, (dim,))),
        {"mu": np.zeros(dim), "kappa": 1.0, "psi": np.eye(dim), "nu": float(dim)},
        param=int(dim))
