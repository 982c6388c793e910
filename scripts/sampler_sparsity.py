import sys, numpy as np
sys.path.insert(0, "/root/repo")
import common_b200 as cb
for wl, nrows in (("C2", 200000), ("C3", 100000), ("C5", 100000)):
    cfg = cb.synth.config(wl)
    k, descs = cfg["k"], cfg["models"]
    arr, z = cb.synth.make_dataset(descs, nrows * 5, k, seed=73, storage=cfg.get("storage"))
    ctx = cb.Context(0)
    st = cb.state(ctx, descs, max_groups=k + 8, cluster_hp={"alpha": 1.0})
    st.bind(cb.numpy_dataview(arr))
    gids = np.asarray([st.create_group() for _ in range(k)])
    st.add_values(gids[z])
    st.sweep(0, nrows, seed=73, sweep=0)
    S = st.read_last_scores()
    x = S - S.max(1, keepdims=True)
    dead = x < -104
    kb = (k + 7) // 8
    pad = np.ones((nrows, kb * 8), bool); pad[:, :k] = dead
    tiles = pad[: nrows // 32 * 32].reshape(nrows // 32, 32, kb, 8)
    print(wl, "elements dead %.3f" % dead.mean(), " (tile,batch) all dead %.3f" % tiles.all(axis=(1, 3)).mean(),
          " rows with >1 live %.3f" % ((~dead).sum(1) > 1).mean())
    st.close(); ctx.close()
