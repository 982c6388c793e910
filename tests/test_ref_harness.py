"""CPU: oracle/_ref (family classes behind the reference's real models/base.hpp + row_accessor)
agrees with the plain-C oracle; skipped where _ref was not built (it needs /root/reference)."""
import numpy as np
import pytest

import common_b200 as cb
import oracle_lib as ol
import ref_lib


@pytest.fixture(scope="module")
def ref():
    r = ref_lib.load()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference; run `make -C oracle ref`)")
    return r


@pytest.mark.parametrize("mask_frac", [0.0, 0.1])
def test_reference_api_loop_matches_c_oracle(oracle, ref, mask_frac):
    descs = [cb.bb, cb.gp, cb.nich, cb.dd(11), cb.niw(3)]
    arr, z = cb.synth.make_dataset(descs, 120, 5, seed=2, mask_frac=mask_frac)
    view = cb.numpy_dataview(arr)
    hp = np.concatenate([oracle.flat_hp(d) for d in descs])
    ss, counts = ol.build_suffstats(oracle, descs, hp, view, z, 5, prec=32)
    lp = ol.logprior(counts, 1.0)
    a = ref.score_rows(descs, hp, ss, lp, view, nthreads=3)
    b = oracle.score_rows(descs, hp, ss, lp, view, f32=True)
    c = oracle.score_rows(descs, hp, ss, lp, view)
    assert np.max(np.abs(a - b) / np.maximum(1, np.abs(b))) < 2e-6   # two float restatements, same op order
    assert np.max(np.abs(a - c) / np.maximum(1, np.abs(c))) < 2e-4   # float suffstats (Welford) + float formulas vs fp64


def test_reference_api_any_storage_type(oracle, ref):
    # value_accessor::get<T> + runtime_cast through the real headers (runtime_value.hpp:46-54)
    descs = [cb.dd(9), cb.gp, cb.nich, cb.bb]
    for storage in (np.uint8, np.int64, np.float64):
        arr, z = cb.synth.make_dataset(descs, 50, 3, seed=4, storage=[storage] * 4)
        view = cb.numpy_dataview(arr)
        hp = np.concatenate([oracle.flat_hp(d) for d in descs])
        ss, counts = ol.build_suffstats(oracle, descs, hp, view, z, 3)
        lp = ol.logprior(counts, 1.0)
        a = ref.score_rows(descs, hp, ss, lp, view)
        c = oracle.score_rows(descs, hp, ss, lp, view)
        assert np.max(np.abs(a - c) / np.maximum(1, np.abs(c))) < 2e-4  # float formulas (gp lgammaf cancellation) vs fp64


def test_perf_group_loop_runs(ref):
    ns, score = ref.perf_group(ol.BB, 0, D=1000, niters=200)   # bin/perf_group.cpp:76-125
    assert ns > 0 and np.isfinite(score) and score < 0


def test_reference_api_loop_with_count_vector_fields(oracle, ref):
    # dm (src/models/dm.cpp) behind the reference's real base.hpp: value_accessor::get<unsigned>(i) over a sub-array
    # field (runtime_value.hpp:46-54), row_accessor::bump over type.n() mask bytes (recarray/dataview.hpp:58-66)
    descs = [cb.dm(6), cb.bb, cb.dm(3)]
    for mask_frac in (0.0, 0.1):
        arr, z = cb.synth.make_dataset(descs, 90, 4, seed=6, mask_frac=mask_frac)
        view = cb.numpy_dataview(arr)
        hp = np.concatenate([oracle.flat_hp(d) for d in descs])
        ss, counts = ol.build_suffstats(oracle, descs, hp, view, z, 4)
        lp = ol.logprior(counts, 1.0)
        a = ref.score_rows(descs, hp, ss, lp, view, nthreads=2)
        b = oracle.score_rows(descs, hp, ss, lp, view, f32=True)
        c = oracle.score_rows(descs, hp, ss, lp, view)
        assert np.max(np.abs(a - b) / np.maximum(1, np.abs(b))) < 2e-6   # two float restatements of dm.cpp:38-76
        # the float formula itself: lgammaf(e + x) - lgammaf(e) cancels terms of size e log e per category (e in the hundreds)
        assert np.max(np.abs(a - c) / np.maximum(1, np.abs(c))) < 1e-3
