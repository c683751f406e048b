"""GT order matrices from InstaOrder string annotations: instaorder_b200.annotations against the UNMODIFIED reference
``InstaOrderDataset.get_gt_ordering`` (datasets/reader.py:335-400) on random annotation dicts (build container only:
needs /root/reference), plus hand-checked cases that run everywhere."""
import numpy as np
import pytest

from instaorder_b200 import annotations as A
from oracle import ref_shim


def _random_ann(rng, n):
    occ, depth = [], []
    for i in range(n):
        for j in range(i + 1, n):
            r = rng.rand()
            if r < 0.25:
                occ.append({"order": "%d<%d" % ((i, j) if rng.rand() < 0.5 else (j, i))})
            elif r < 0.35:
                occ.append({"order": "%d<%d & %d<%d" % (i, j, j, i)})
            if rng.rand() < 0.7:
                a, b = (i, j) if rng.rand() < 0.5 else (j, i)
                depth.append({"order": ("%d<%d" if rng.rand() < 0.8 else "%d=%d") % (a, b),
                              "overlap": bool(rng.rand() < 0.3), "count": int(rng.randint(1, 4))})
    return {"instance_ids": list(range(n)), "occlusion": occ, "depth": depth}


def test_hand_checked():
    ann = {"instance_ids": [0, 1, 2], "occlusion": [{"order": "0<1"}, {"order": "1<2 & 2<1"}],
           "depth": [{"order": "2<0", "overlap": True, "count": 2}, {"order": "0=1", "overlap": False, "count": 3}]}
    occ = A.gt_ordering(ann, "occlusion")
    assert occ.tolist() == [[0, 1, 0], [0, 0, 1], [0, 1, 0]]
    d, ov, cnt = A.gt_ordering(ann, "depth")
    assert d.tolist() == [[-1, 2, 0], [2, -1, -1], [1, -1, -1]]
    assert ov.tolist() == [[-1, 0, 1], [0, -1, -1], [1, -1, -1]]
    assert cnt.tolist() == [[-1, 3, 2], [3, -1, -1], [2, -1, -1]]
    assert A.gt_ordering(ann, "depth", rm_overlap=1)[1].tolist() == [[-1, 0, -1], [0, -1, -1], [-1, -1, -1]]
    assert A.gt_ordering(ann, "occlusion", rm_bidirec=1).tolist() == [[0, -1, 0], [-1, 0, 0], [0, 0, 0]]   # the quirk


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_against_reference_reader():
    from oracle import gen_golden_traindata as GG
    r_datasets = GG.load_reference_datasets()
    reader = r_datasets.reader.InstaOrderDataset
    rng = np.random.RandomState(0)
    for trial in range(40):
        ann = _random_ann(rng, int(rng.randint(2, 8)))
        ref = object.__new__(reader)
        ref.annot_info = [ann]
        for rm in (0, 1):
            want = ref.get_gt_ordering(0, "depth", rm_overlap=rm)
            got = A.gt_ordering(ann, "depth", rm_overlap=rm)
            for w, g in zip(want, got):
                assert np.array_equal(np.asarray(w), g)
        assert np.array_equal(np.asarray(ref.get_gt_ordering(0, "occlusion")), A.gt_ordering(ann, "occlusion"))
        try:
            want = np.asarray(ref.get_gt_ordering(0, "occlusion", rm_bidirec=1))
        except UnboundLocalError:
            with pytest.raises(UnboundLocalError):
                A.gt_ordering(ann, "occlusion", rm_bidirec=1)
        else:
            assert np.array_equal(want, A.gt_ordering(ann, "occlusion", rm_bidirec=1))
