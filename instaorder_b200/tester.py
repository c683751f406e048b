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
        ``input_size``, ``remove_occ_bidirec``, ``remove_depth_overlap``, ``use_category``, ``enlarge_box``,
        ``dataset``, ``trainval_dataset``) as in the reference's yaml + CLI; model: an
        ``instaorder_b200.models`` wrapper in eval mode (``None`` for the 'area' / 'yaxis' heuristics); data_reader: see
        the module docstring;
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

    _KIND = {"SupDepthOrderDataset": "depth", "SupOcclusionOrderDataset": "occ", "PartialCompDataset": "occ",
             "SupDepthOccOrderDataset": "od"}                                    # tools/test.py:165-174
    _METHODS = {"od": ("InstaOrderNet_od", "InstaDepthNet_od"),                  # tools/test.py:207, :309-334, :421-447
                "depth": ("area", "yaxis", "InstaOrderNet_d", "midas_pretrained", "InstaDepthNet_d"),
                "occ": ("area", "yaxis", "InstaOrderNet_o", "OrderNet")}

    def _kind(self):
        a = self.args
        kind = self._KIND.get(a.data.get("trainval_dataset"))
        if kind is None:        # no dataset class named: the method decides (the heuristics need the dataset class)
            kind = {m: k for k in ("od", "depth", "occ") for m in self._METHODS[k] if m not in ("area", "yaxis")}.get(
                a.order_method)
        if kind is None or a.order_method not in self._METHODS[kind]:
            raise Exception("No such order method: {}".format(a.order_method))      # tools/test.py:221, :336, :449
        return kind

    def _heuristic(self, kind, modal):
        """tools/test.py:309-322 (depth) / :421-434 (occlusion): which instance wins depends on the dataset."""
        a = self.args
        coco_like = a.data.get("dataset", "InstaOrder") in ("COCOA", "InstaOrder")
        m = modal.cpu().numpy() if hasattr(modal, "is_cuda") else modal
        if kind == "depth":
            if a.order_method == "area":
                return _infer.infer_depth_order_area(m, closer="larger")
            return _infer.infer_depth_order_yaxis(m, closer="lower" if coco_like else "higher")
        if a.order_method == "area":
            return _infer.infer_occ_order_area(m, occluder="larger")
        return _infer.infer_occ_order_yaxis(m, occluder="lower" if coco_like else "higher")

    def run(self, indices=None):
        """Evaluates images ``indices`` (default: this rank's round-robin share of ``len(data_reader)``) and returns the
        dict of dataset-level numbers (all ranks get the same dict)."""
        import torch.distributed as dist
        a = self.args
        method = a.order_method
        kind = self._kind()
        dataset = a.data.get("dataset", "InstaOrder")
        n_total = len(self.reader)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        mine = list(indices) if indices is not None else _sharding.shard_interleaved(n_total, rank, world)
        want_occ, want_depth = kind in ("occ", "od"), kind in ("depth", "od")
        dsm = getattr(a, "disp_select_method", "")
        # the disparity-derived depth orders (reference inference.py:576-599) are per-image calls; everything else
        # that runs a network goes through the batched engine; the heuristics need no network at all
        per_image = kind == "depth" and (method == "midas_pretrained" or (method == "InstaDepthNet_d" and dsm != ""))
        batched = method not in ("area", "yaxis") and not per_image
        eng = self.model.engine_for(a.data["input_size"]) if batched else None
        prf_rows, whdr_rows = [], []
        inflight = []

        def finalize(entry):
            res, gts_occ, gts_depth, names = entry
            if batched:
                if not isinstance(res, list):
                    res = eng.collect(res)
                pred_occ = [r["occ"] for r in res] if want_occ else None
                pred_depth = [r["depth"] for r in res] if want_depth else None
            else:
                pred_occ, pred_depth = (res if want_occ else None), (res if want_depth else None)
            if want_occ:
                rows = _engine.metrics_prf(pred_occ, gts_occ, a.zd)
                prf_rows.extend(rows.tolist())
            if want_depth:
                rows = _engine.metrics_whdr(pred_depth, [g[0] for g in gts_depth], [g[1] for g in gts_depth],
                                            [g[2] for g in gts_depth])
                whdr_rows.extend(rows.tolist())
            for k, fn in enumerate(names):
                if want_depth:
                    w = whdr_rows[len(whdr_rows) - len(names) + k]
                    self._log("[%s]\t%.3f | %.3f | %.3f" % (fn, w[2], w[5], w[8]))      # ovlX_all | ovlO_all | ovlOX_all
                if want_occ:
                    p = prf_rows[len(prf_rows) - len(names) + k]
                    self._log("\t\t\trecall=%.3f / precision=%.3f / f1=%.3f" % (p[0], p[1], p[2]))

        for c0 in range(0, len(mine), self.images_per_call):
            chunk = mine[c0:c0 + self.images_per_call]
            scenes, gts_occ, gts_depth, names, preds = [], [], [], [], []
            for i in chunk:
                modal, category, bboxes, amodal_gt, image_fn = self.reader.get_image_instances(i, with_gt=True)[:5]
                if kind != "od" and a.data.get("use_category", False):            # tools/test.py:294-295, :405-406
                    if hasattr(modal, "is_cuda"):      # masks left in HBM by the reader (device_masks=True)
                        import torch
                        modal = modal * torch.as_tensor(np.asarray(category), dtype=modal.dtype,
                                                        device=modal.device)[:, None, None]
                    else:
                        modal = modal * np.asarray(category)[:, None, None]
                names.append(image_fn)
                if want_depth:                                                     # tools/test.py:201-203, :300-303
                    gts_depth.append(self.reader.get_gt_ordering(i, "depth") if kind == "od" else
                                     self.reader.get_gt_ordering(i, "depth",
                                                                 rm_overlap=a.data.get("remove_depth_overlap", 0)))
                if want_occ:                                                       # tools/test.py:204, :413-418
                    if dataset == "InstaOrder":
                        g = self.reader.get_gt_ordering(i, "occlusion", a.data.get("remove_occ_bidirec", 0))
                    elif getattr(self.reader, "dataset", dataset) == "KINS":       # gt_ordering == "man"
                        g = _infer.infer_gt_order(modal.cpu().numpy() if hasattr(modal, "is_cuda") else modal, amodal_gt)
                    else:
                        g = self.reader.get_gt_ordering(i)
                    if not np.any(np.asarray(g) != -1):
                        raise ValueError("image %r: no occlusion ground truth entry != -1 (the reference's sklearn "
                                         "call raises on the empty selection)" % (image_fn,))
                    gts_occ.append(g)
                if batched:
                    scenes.append(_engine.Scene(self.load_image(image_fn), modal, self.expand_bbox(bboxes)))
                elif per_image:
                    preds.append(_infer.infer_order_sup_depth(
                        self.model, self.load_image(image_fn), modal, self.expand_bbox(bboxes), a.pairs, method,
                        a.data["patch_or_image"], a.data["input_size"], dsm,
                        use_rgb=(getattr(a, "model", None) or {}).get("use_rgb", True))[0])
                else:
                    preds.append(self._heuristic(kind, modal))
            # the batched engine runs one chunk behind: while its kernels work on this chunk, the loop above loads (and
            # rasterises) the next one; results are collected, scored and logged in image order
            if batched and hasattr(eng, "submit_scenes"):
                entry = (eng.submit_scenes(scenes, method, pairs=a.pairs, patch_or_image=a.data["patch_or_image"]),
                         gts_occ, gts_depth, names)
            elif batched:
                entry = (eng.infer_scenes(scenes, method, pairs=a.pairs, patch_or_image=a.data["patch_or_image"]),
                         gts_occ, gts_depth, names)
            else:
                entry = (preds, gts_occ, gts_depth, names)
            inflight.append(entry)
            if len(inflight) > 1:
                finalize(inflight.pop(0))
        while inflight:
            finalize(inflight.pop(0))
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
