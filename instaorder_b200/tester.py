"""Evaluation driver with the reference's structure (``tools/test.py::Tester``, :136-286 / :287-400 / :402-482):
``prepare_model`` -> per-image order inference -> per-image P/R/F1 and WHDR -> the dataset-level numbers under the
reference's wandb keys (``val_{ovl}/WHDR_{eq}``, ``val/recall``, ``val/precision``, ``val/f1``, ``val/num_test_images``).

Differences from the reference loop, none visible in the numbers: the data reader is INJECTED (any object with the
reference reader's ``get_image_instances(i, with_gt=True)`` / ``get_gt_ordering(i, kind[, rm_bidirec])`` methods --
file / annotation I/O stays out of scope; ``instaorder_b200.masks`` / ``annotations`` provide the GPU mask
and GT-matrix producers for such a reader), images are processed ``images_per_call`` at a time through the batched
engine instead of one ``infer_order_sup_*`` call each, the metrics of a call run as two batched kernel launches, and with
``torch.distributed`` initialised the images are sharded round-robin over the ranks (no collective in the loop, one
gather of the per-image rows at the end).  PNG / graph dumps (``--save_pngs``) are not reproduced."""
import numpy as np

from . import engine as _engine, inference as _infer, sharding as _sharding


class Tester(object):
    def __init__(self, args, model, data_reader, load_image, logger=None, wb_logger=None):
        """args: ``order_method``, ``pairs``, ``zd``, ``disp_select_method``, ``data`` = dict(``patch_or_image``,
        ``input_size``, ``remove_occ_bidirec``, ``use_category``, ``enlarge_box``) as in the reference's yaml + CLI;
        model: an ``instaorder_b200.models`` wrapper in eval mode; data_reader: see the module docstring;
        load_image(image_fn) -> uint8 [H, W, 3]."""
        self.args, self.model, self.reader, self.load_image = args, model, data_reader, load_image
        self.logger, self.wb_logger = logger, wb_logger
        self.images_per_call = int(getattr(args, "images_per_call", 16))
        self.curr_step = int(getattr(args, "curr_step", 0))

    # tools/test.py:155-163
    def expand_bbox(self, bboxes):
        return _engine.expand_bbox(bboxes, float(self.args.data.get("enlarge_box", 3.0)))

    def _log(self, msg):
        if self.logger is not None:
            self.logger.info(msg)

    def run(self, indices=None):
        """Evaluates images ``indices`` (default: this rank's round-robin share of ``len(data_reader)``) and returns the
        dict of dataset-level numbers (all ranks get the same dict)."""
        import torch.distributed as dist
        a = self.args
        method = a.order_method
        n_total = len(self.reader)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        mine = list(indices) if indices is not None else _sharding.shard_interleaved(n_total, rank, world)
        want_occ = method in ("OrderNet", "InstaOrderNet_o", "InstaOrderNet_od", "InstaDepthNet_od")
        want_depth = method in ("InstaOrderNet_d", "InstaOrderNet_od", "InstaDepthNet_od", "InstaDepthNet_d")
        if not (want_occ or want_depth):
            raise Exception("No such order method: {}".format(method))      # tools/test.py:221
        eng = self.model.engine_for(a.data["input_size"])
        prf_rows, whdr_rows = [], []
        for c0 in range(0, len(mine), self.images_per_call):
            chunk = mine[c0:c0 + self.images_per_call]
            scenes, gts_occ, gts_depth, names = [], [], [], []
            for i in chunk:
                modal, category, bboxes, _, image_fn = self.reader.get_image_instances(i, with_gt=True)[:5]
                if a.data.get("use_category", False):                             # tools/test.py:302-303
                    if hasattr(modal, "is_cuda"):      # masks left in HBM by the reader (device_masks=True)
                        import torch
                        modal = modal * torch.as_tensor(np.asarray(category), dtype=modal.dtype,
                                                        device=modal.device)[:, None, None]
                    else:
                        modal = modal * np.asarray(category)[:, None, None]
                scenes.append(_engine.Scene(self.load_image(image_fn), modal, self.expand_bbox(bboxes)))
                names.append(image_fn)
                if want_depth:
                    gts_depth.append(self.reader.get_gt_ordering(i, "depth"))
                if want_occ:
                    gts_occ.append(self.reader.get_gt_ordering(i, "occlusion", a.data.get("remove_occ_bidirec", 0)))
            res = eng.infer_scenes(scenes, method, pairs=a.pairs, patch_or_image=a.data["patch_or_image"])
            if want_occ:
                rows = _engine.metrics_prf([r["occ"] for r in res], gts_occ, a.zd, device=str(eng.device))
                prf_rows.extend(rows.tolist())
            if want_depth:
                rows = _engine.metrics_whdr([r["depth"] for r in res], [g[0] for g in gts_depth],
                                            [g[1] for g in gts_depth], [g[2] for g in gts_depth], device=str(eng.device))
                whdr_rows.extend(rows.tolist())
            for k, fn in enumerate(names):
                if want_depth:
                    w = whdr_rows[len(whdr_rows) - len(names) + k]
                    self._log("[%s]\t%.3f | %.3f | %.3f" % (fn, w[2], w[5], w[8]))      # ovlX_all | ovlO_all | ovlOX_all
                if want_occ:
                    p = prf_rows[len(prf_rows) - len(names) + k]
                    self._log("\t\t\trecall=%.3f / precision=%.3f / f1=%.3f" % (p[0], p[1], p[2]))
        if world > 1:
            prf = _sharding.gather_metric_rows(np.asarray(prf_rows, np.float64).reshape(-1, 3), mine, n_total) if want_occ else None
            whdr = _sharding.gather_metric_rows(np.asarray(whdr_rows, np.float64).reshape(-1, 9), mine, n_total) if want_depth else None
        else:
            prf, whdr = (prf_rows if want_occ else None), (whdr_rows if want_depth else None)
        agg = _sharding.aggregate_metrics(prf, whdr)
        out = {}
        for key in _infer.WHDR_KEYS:                                      # tools/test.py:264-271
            if "WHDR_" + key in agg:
                ovl, eq = key.split("_")
                out["val_%s/WHDR_%s" % (ovl, eq)] = agg["WHDR_" + key]
                self._log("%s: %s" % (key, agg["WHDR_" + key]))
        if "recall" in agg:                                               # tools/test.py:273-283
            out.update({"val/recall": agg["recall"], "val/precision": agg["precision"], "val/f1": agg["f1"]})
            self._log("\n\n[AVERAGE] recall=%.3f / precision=%.3f / f1=%.3f" % (agg["recall"], agg["precision"], agg["f1"]))
        out["val/num_test_images"] = len(mine) if world == 1 else n_total
        out["val/iter"] = self.curr_step
        if self.wb_logger is not None and rank == 0:
            self.wb_logger.log(out, step=self.curr_step)
        return out
