"""``gemlib.mcmc.Posterior``: the on-disk posterior every downstream step of the reference consumes
(constructed inference.py:352-358; ``write_samples`` / ``write_results`` :376-380 etc.; read back :588-606 and
by ``thin_posterior``, covid19uk/posterior/thin.py:7-21).

Layout: ``samples/<key>`` and ``results/<nested/key>`` datasets of shape
``[num_samples, ...]`` created from example dictionaries, filled burst by burst along the first axis.
Device tensors are copied to the host here (one D2H per burst and dataset).
"""
from __future__ import annotations

import numpy as np

from .. import hdf5_min


def _to_numpy(x):
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.asarray(x)


def _flatten(tree, prefix=""):
    for key, value in tree.items():
        path = f"{prefix}/{key}" if prefix else str(key)
        if isinstance(value, dict):
            yield from _flatten(value, path)
        else:
            yield path, value


class Posterior:
    def __init__(self, filename, sample_dict, results_dict, num_samples):
        self._file = hdf5_min.File(filename, "w")
        self.num_samples = int(num_samples)
        for root, tree in (("samples", sample_dict), ("results", results_dict)):
            for path, example in _flatten(tree):
                ex = _to_numpy(example)
                self._file.create_dataset(f"{root}/{path}", shape=(self.num_samples,) + tuple(ex.shape[1:]), dtype=ex.dtype)

    def _write(self, root, tree, first_dim_offset):
        off = int(first_dim_offset)
        for path, value in _flatten(tree):
            arr = _to_numpy(value)
            self._file[f"{root}/{path}"][off: off + arr.shape[0]] = arr

    def write_samples(self, draws_dict, first_dim_offset=0):
        self._write("samples", draws_dict, first_dim_offset)

    def write_results(self, results_dict, first_dim_offset=0):
        self._write("results", results_dict, first_dim_offset)

    def __getitem__(self, path):
        return self._file[path]

    def close(self):
        if self._file is not None:
            self._file.close()
            self._file = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass


def thin_posterior(input_file, output_file, config):
    """covid19uk/posterior/thin.py:7-21: every ``config["by"]``-th sample of ``samples/*`` in
    [``start``, ``end``) plus ``initial_state``, pickled to ``output_file``; the dictionary is also returned."""
    import pickle

    idx = slice(config["start"], config["end"], config["by"])
    f = hdf5_min.File(input_file, "r")
    out = {key: f[f"samples/{key}"][idx] for key in f["samples"].keys()}
    if "initial_state" in f:
        out["initial_state"] = f["initial_state"][:]
    f.close()
    if output_file is not None:
        with open(output_file, "wb") as fh:
            pickle.dump(out, fh)
    return out
