import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import common_b200 as cb, oracle_lib as ol
orc = ol.load(); ctx = cb.Context(0)
for dim in (64, 5):
    for off in (0.0, 10.0, 100.0, 1000.0):
        descs = [cb.niw(dim)]
        n, k = 2000, 6
        arr, z = cb.synth.make_dataset(descs, n, k, seed=3)
        data = np.array(arr, copy=True)
        nm = data.dtype.names[0]
        data[nm] = data[nm] + off
        view = cb.numpy_dataview(data)
        st = cb.state(ctx, descs, max_groups=k + 2, cluster_hp={"alpha": 1.0})
        st.bind(view)
        gids = [st.create_group() for _ in range(k)]
        st.add_values(np.asarray(gids)[z])
        hp = np.concatenate([orc.flat_hp(d) for d in descs])
        ss, counts = ol.build_suffstats(orc, descs, hp, view, z, k)
        want = orc.score_rows(descs, hp, ss, ol.logprior(counts, 1.0), view)
        _, S = st.score_rows()
        _, S64 = st.score_rows_f64()
        e = lambda a: np.max(np.abs(a - want) / np.maximum(1, np.abs(want)))
        print("dim %2d offset %7.1f  fp32 path rel err %.2e   fp64 path %.2e" % (dim, off, e(S), e(S64)))
        st.close()
