"""Golden vectors produced by the REFERENCE'S OWN CODE (run in the build container, where /root/reference exists):

* ``get_mosaic_of_centers`` and ``split_cloud_into_samples`` of myria3d/pctl/dataset/utils.py:29-38,126-158 -- the module
  is loaded from its file with an empty stand-in for the uninstalled ``pdal`` import, and its LAS reader
  (``pdal_read_las_array_as_float32``) replaced by a function returning seeded synthetic points; the splitting code that
  runs is the reference's, unmodified;
* ``subsample_data`` / ``MinimumNumNodes`` / ``MaximumNumNodes`` / ``NormalizePos`` of myria3d/pctl/transforms/transforms.py:30-84,
  149-165 -- loaded the same way, with a 20-line stand-in for ``torch_geometric.data.Data`` (attribute bag, iteration over
  (key, value), ``num_nodes``) and ``BaseTransform = object``; torch's CPU generator is seeded, so ``randperm`` is the
  reference's stream.

    python oracle/gen_golden_ref.py        ->  tests/golden/ref_sample_prep.npz  (~260 KB)

TEST INFRASTRUCTURE: nothing in the product imports this; the GPU box never runs it (no /root/reference there)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_sample_prep.npz")


class _Data:  # the few torch_geometric.data.Data behaviours transforms.py relies on
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __iter__(self):
        return iter(list(self.__dict__.items()))

    def __getitem__(self, k):
        return self.__dict__[k]

    def __setitem__(self, k, v):
        self.__dict__[k] = v

    @property
    def keys(self):
        return list(self.__dict__.keys())


def _load(path, name, stubs):
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main():
    class _Anything(types.ModuleType):  # annotations such as pdal.Reader.las are evaluated at import time
        def __getattr__(self, name):
            return _Anything(name)

    pdal = _Anything("pdal")
    utils = _load(os.path.join(REF, "myria3d/pctl/dataset/utils.py"), "ref_dataset_utils", {"pdal": pdal})

    tg, tgd, tgt = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data"), types.ModuleType("torch_geometric.transforms")
    tgd.Data, tgt.BaseTransform = _Data, object
    m3, m3u, m3uu = types.ModuleType("myria3d"), types.ModuleType("myria3d.utils"), types.ModuleType("myria3d.utils.utils")
    import logging
    m3uu.get_logger = logging.getLogger
    m3u.utils = m3uu
    tr = _load(os.path.join(REF, "myria3d/pctl/transforms/transforms.py"), "ref_transforms",
               {"torch_geometric": tg, "torch_geometric.data": tgd, "torch_geometric.transforms": tgt,
                "myria3d": m3, "myria3d.utils": m3u, "myria3d.utils.utils": m3uu})

    out = {}
    # ---- receptive-field split: 3 layouts (no overlap; the predict overlap of 25 m; a ragged tile width)
    for tag, (tile, sub, ov, n, seed) in {"a": (200, 50, 0, 6000, 1), "b": (200, 50, 25, 6000, 2), "c": (110, 50, 10, 2500, 3)}.items():
        g = np.random.default_rng(seed)
        pts = np.zeros(n, dtype=[("X", "f4"), ("Y", "f4"), ("Z", "f4")])
        pts["X"] = (g.random(n) * tile + 1000.0).astype(np.float32)
        pts["Y"] = (g.random(n) * tile + 2000.0).astype(np.float32)
        pts["Z"] = (g.random(n) * 30.0).astype(np.float32)
        pts["X"][:40] = pts["X"].min() + np.arange(40, dtype=np.float32) * np.float32(tile / 40.0)  # points on field borders
        utils.pdal_read_las_array_as_float32 = lambda path, epsg, _p=pts: _p
        samples = [np.sort(idx) for idx, _ in utils.split_cloud_into_samples("unused.las", tile, sub, "2154", subtile_overlap=ov)]
        out[f"split_{tag}_args"] = np.array([tile, sub, ov], dtype=np.float64)
        out[f"split_{tag}_pos"] = np.stack([pts["X"], pts["Y"], pts["Z"]], axis=1)
        out[f"split_{tag}_idx"] = np.concatenate(samples).astype(np.int32)
        out[f"split_{tag}_off"] = np.cumsum([0] + [len(s) for s in samples]).astype(np.int64)
        out[f"split_{tag}_centers"] = np.stack(utils.get_mosaic_of_centers(tile, sub, subtile_overlap=ov))
    out["mosaic_counts"] = np.array([len(utils.get_mosaic_of_centers(1000, 50, 0)), len(utils.get_mosaic_of_centers(1000, 50, 25))])

    # ---- node budgets and NormalizePos
    torch.manual_seed(2024)
    pos = torch.rand(500, 3) * 50.0 - 25.0
    x = torch.rand(500, 4)
    d = tr.MaximumNumNodes(200)(_Data(pos=pos.clone(), x=x.clone(), num_nodes=500, idx_in_original_cloud=np.arange(500)))
    out["max_pos_in"], out["max_x_in"] = pos.numpy(), x.numpy()
    out["max_pos_out"], out["max_x_out"], out["max_num_nodes"] = d.pos.numpy(), d.x.numpy(), np.array(int(d.num_nodes))
    torch.manual_seed(2025)
    d = tr.MinimumNumNodes(300)(_Data(pos=pos[:70].clone(), x=x[:70].clone(), num_nodes=70))
    out["min_pos_out"], out["min_x_out"], out["min_num_nodes"] = d.pos.numpy(), d.x.numpy(), np.array(int(d.num_nodes))
    d = tr.NormalizePos(subtile_width=50)(_Data(pos=pos.clone()))
    out["normalize_pos_out"] = d.pos.numpy()
    # ---- sliding-window reduction of overlapping predictions: Interpolator.store_predictions + reduce_predicted_logits
    # (myria3d/models/interpolation.py:94-121).  torch_scatter is not installable: its documented CPU behaviour for
    # scatter_sum(src, index, out=out, dim=0) -- rows added to out[index] in input order -- is supplied as index_add_.
    ts = types.ModuleType("torch_scatter")

    def scatter_sum(src, index, dim=0, out=None):
        assert dim == 0 and out is not None
        return out.index_add_(0, index, src)

    ts.scatter_sum = scatter_sum
    interp = _load(os.path.join(REF, "myria3d/models/interpolation.py"), "ref_interpolation",
                   {"pdal": pdal, "torch_scatter": ts, "pdaltools": _Anything("pdaltools"),
                    "myria3d": _Anything("myria3d"), "myria3d.pctl": _Anything("myria3d.pctl"),
                    "myria3d.pctl.dataset": _Anything("myria3d.pctl.dataset"),
                    "myria3d.pctl.dataset.utils": _Anything("myria3d.pctl.dataset.utils")})
    it = interp.Interpolator(interpolation_k=10, classification_dict={1: "a", 2: "b", 6: "c", 9: "d"})
    g = torch.Generator().manual_seed(77)
    nb_points, logit_batches, idx_batches = 3000, [], []
    for b in range(3):  # 3 predict batches of 2 windows each; neighbouring windows overlap
        lg, ids = [], []
        for w in range(2):
            start = (2 * b + w) * 400
            idx = torch.arange(start, start + 900)[torch.randperm(900, generator=g)[:700]].numpy()
            ids.append(idx)
            lg.append(torch.randn(700, 4, generator=g))
        logit_batches.append(torch.cat(lg))
        idx_batches.append(ids)
        it.store_predictions(logit_batches[-1], ids)
    reduced, idx_full = it.reduce_predicted_logits(nb_points)
    out["stitch_logits"] = torch.cat(logit_batches).numpy()
    out["stitch_idx"] = np.concatenate([np.concatenate(ids) for ids in idx_batches]).astype(np.int64)
    out["stitch_nb_points"] = np.array(nb_points)
    out["stitch_reduced"], out["stitch_idx_out"] = reduced.numpy(), np.asarray(idx_full).astype(np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", {k: v.shape for k, v in out.items() if k.startswith("split_") and k.endswith("_off")})


if __name__ == "__main__":
    main()
