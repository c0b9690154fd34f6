"""Per-cloud prediction files: the hand-off between the network half and the pose half of the hot path.

Mirror of lib/prediction_io.py:65-95 (`save_batch_nn`): same function name, argument order, dataset names, shapes and
dtypes, one HDF5 file per cloud named `<basename>.h5`.  The reference writes through h5py; where h5py is importable it
is used, otherwise (this image) the files are written by `minih5` -- real HDF5 (superblock v2, contiguous datasets in
the root group), which the unmodified evaluation scripts open with `h5py.File`.  `load_prediction` reads them back
(h5py, else minih5) and returns a mapping that indexes like the reference's `h5py.File`
(`f['nocs_per_point'][idx, 3*j:3*(j+1)]`, evaluation/parallel_ancsh_pose.py:224-260); `.npz` files of earlier versions
of this package are still readable.

`PredictionStore` is the in-memory version of the same hand-off: `save_batch_nn(..., save_dir=store)`
keeps the per-cloud dicts in RAM so the pose stage can consume them without touching the disk.  The fully fused path
(`pipeline.AncshPipeline`) does not even leave HBM; these formats exist for drop-in use with the reference's scripts.
"""
import os

import numpy as np

from . import minih5

try:  # pragma: no cover - h5py is absent in the build image
    import h5py
except Exception:  # noqa: BLE001
    h5py = None

# datasets of one per-cloud file, in the order lib/prediction_io.py:73-92 creates them
DATASETS = ("confidence_per_point", "P", "cls_gt", "nocs_gt", "nocs_per_point", "instance_per_point", "gocs_per_point",
            "nocs_gt_g", "heatmap_per_point", "heatmap_gt", "unitvec_gt", "unitvec_per_point", "joint_axis_per_point",
            "joint_axis_gt", "index_per_point", "joint_cls_gt")


class PredictionStore(dict):
    """basename -> {dataset name: array, 'attrs': {...}}; pass as `save_dir` to keep predictions in memory."""


def _records(nn_name, pred_result, input_batch, basename_list, is_mixed, W_reduced, two_stages):
    batch_size = pred_result["W"].shape[0]
    assert batch_size == len(basename_list), \
        "Oh no, batch size is {}, while len of basename_list is{}".format(batch_size, len(basename_list))
    instance_per_point = pred_result["W"]                                  # BxNxK
    if W_reduced:
        instance_per_point = np.argmax(instance_per_point, axis=2)         # BxN (prediction_io.py:70-71)
    for b in range(batch_size):
        d = {
            "confidence_per_point": pred_result["confi_per_point"][b],
            "P": input_batch["P"][b],
            "cls_gt": input_batch["cls_gt"][b],
            "nocs_gt": input_batch["nocs_gt"][b],
            "nocs_per_point": pred_result["nocs_per_point"][b],
            "instance_per_point": instance_per_point[b],
        }
        if is_mixed:
            d["gocs_per_point"] = pred_result["gocs_per_point"][b]
        d.update({
            "nocs_gt_g": input_batch["nocs_gt_g"][b],
            "heatmap_per_point": pred_result["heatmap_per_point"][b],
            "heatmap_gt": input_batch["heatmap_gt"][b],
            "unitvec_gt": input_batch["unitvec_gt"][b],
            "unitvec_per_point": pred_result["unitvec_per_point"][b],
            "joint_axis_per_point": pred_result["joint_axis_per_point"][b],
            "joint_axis_gt": input_batch["orient_gt"][b],
            "index_per_point": pred_result["index_per_point"][b],
            "joint_cls_gt": input_batch["joint_cls_gt"][b],
        })
        if two_stages:
            d["joint_params_pred"] = pred_result["joint_params_pred"][b]
            d["joint_params_gt"] = input_batch["joint_params_gt"][b]
        yield basename_list[b], {"method_name": nn_name, "basename": basename_list[b]}, d


def save_batch_nn(nn_name, pred_result, input_batch, basename_list, save_dir, sample_index=None, is_mixed=False,
                  W_reduced=True, two_stages=False):
    """lib/prediction_io.py:65-95.  `save_dir`: directory (one `<basename>.h5` per cloud) or a PredictionStore."""
    for base, attrs, d in _records(nn_name, pred_result, input_batch, basename_list, is_mixed, W_reduced, two_stages):
        if isinstance(save_dir, PredictionStore):
            save_dir[base] = dict({k: np.asarray(v) for k, v in d.items()}, attrs=attrs)
        elif h5py is not None:
            with h5py.File(os.path.join(save_dir, base + ".h5"), "w") as f:
                for k, v in attrs.items():
                    f.attrs[k] = v
                for k, v in d.items():
                    f.create_dataset(k, data=v)
        else:
            minih5.write(os.path.join(save_dir, base + ".h5"), d, attrs)


def load_prediction(save_dir, basename):
    """-> mapping dataset name -> array (h5py.File if the .h5 exists and h5py is importable, else the .npz / store
    entry).  Index it like the reference does: f['P'][idx, :3], f['instance_per_point'][()]."""
    if isinstance(save_dir, PredictionStore):
        return _ArrayFile(save_dir[basename])
    h5 = os.path.join(save_dir, basename + ".h5")
    if os.path.exists(h5):
        if h5py is not None:
            return h5py.File(h5, "r")
        d, attrs = minih5.read(h5)
        return _ArrayFile(dict(d, attrs=attrs))
    z = np.load(os.path.join(save_dir, basename + ".npz"))
    d = {k: z[k] for k in z.files if not k.startswith("attrs/")}
    d["attrs"] = {k[6:]: z[k].item() for k in z.files if k.startswith("attrs/")}
    return _ArrayFile(d)


def list_predictions(save_dir):
    """File names of a prediction directory as `os.listdir` shows them to pose_multi_process.py:36 (`<basename>.h5`)."""
    if isinstance(save_dir, PredictionStore):
        return [b + ".h5" for b in save_dir]
    return sorted(os.path.splitext(f)[0] + ".h5" for f in os.listdir(save_dir) if f.endswith((".h5", ".npz")))


class _Dataset:
    """numpy array that also answers h5py's `ds[()]`."""

    def __init__(self, a):
        self.a = np.asarray(a)
        self.shape, self.dtype = self.a.shape, self.a.dtype

    def __getitem__(self, key):
        return self.a if (isinstance(key, tuple) and len(key) == 0) else self.a[key]

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)


class _ArrayFile:
    def __init__(self, d):
        self.attrs = d.get("attrs", {})
        self._d = {k: v for k, v in d.items() if k != "attrs"}

    def __getitem__(self, k):
        return _Dataset(self._d[k])

    def __contains__(self, k):
        return k in self._d

    def keys(self):
        return self._d.keys()

    def close(self):
        pass
