"""Interchange files of the hot path.

The reference hands data between its CLI stages through HDF5 files written with h5py and
``pandas.to_hdf`` (region_params, sequence_model_192/64, genome_counts, all_window_genome_counts, idx,
mappability, window_{W}/..., pretrained tables; DigPreprocess.py:63-73, DigPretrain.py:82-96,156-177,
207-208).  Neither h5py nor PyTables exists in this image, so the same key layout is served by a
directory store (``<path>`` is a directory holding ``<key>.npy`` / ``<key>.table.npz`` files and
``attrs.json``).  When h5py AND tables are importable and the path is an existing HDF5 file, the real
file is read instead, so pretrained models produced by the reference can be dropped in.
"""
import json
import os

import numpy as np
import pandas as pd


def _have_hdf5():
    try:
        import h5py  # noqa: F401
        import tables  # noqa: F401
        return True
    except Exception:
        return False


def _is_hdf5_file(path):
    if not os.path.isfile(path):
        return False
    with open(path, "rb") as f:
        return f.read(8) == b"\x89HDF\r\n\x1a\n"


def _key_path(path, key, suffix):
    return os.path.join(path, key.strip("/").replace("/", "__") + suffix)


class Store:
    """Directory-backed key/value store with the reference's HDF5 key names."""

    def __init__(self, path, mode="a"):
        self.path = str(path)
        self.hdf5 = _is_hdf5_file(self.path)
        if self.hdf5 and not _have_hdf5():
            raise RuntimeError("%s is an HDF5 file but h5py/PyTables are not installed in this environment; "
                               "convert it with the reference's environment or install them" % self.path)
        if not self.hdf5:
            if mode == "r" and not os.path.isdir(self.path):
                raise FileNotFoundError(self.path)
            if mode == "w" and os.path.isdir(self.path):
                for f in os.listdir(self.path):
                    os.remove(os.path.join(self.path, f))
            os.makedirs(self.path, exist_ok=True)

    # ---- tables (pandas objects)
    def write_table(self, key, df):
        if self.hdf5:
            df.to_hdf(self.path, key=key, mode="a")
            return
        if isinstance(df, pd.Series):
            df = df.to_frame(name="__series__")
        def plain(x):
            v = np.asarray(x)
            if v.dtype.kind in "OUST":              # strings (object / pandas string dtype) -> fixed-width unicode
                v = np.array([str(e) for e in v], dtype=str) if len(v) else np.zeros(0, dtype="U1")
            return v
        cols = {"col%d" % i: plain(df[c]) for i, c in enumerate(df.columns)}
        np.savez(_key_path(self.path, key, ".table.npz"), __columns__=np.array([str(c) for c in df.columns]),
                 __index__=plain(df.index), __index_name__=np.array([str(df.index.name or "")]), **cols)

    def read_table(self, key):
        if self.hdf5:
            return pd.read_hdf(self.path, key)
        z = np.load(_key_path(self.path, key, ".table.npz"), allow_pickle=False)
        cols = [str(c) for c in z["__columns__"]]
        df = pd.DataFrame({c: z["col%d" % i] for i, c in enumerate(cols)}, index=z["__index__"])
        name = str(z["__index_name__"][0])
        df.index.name = name or None
        if cols == ["__series__"]:
            return df["__series__"].rename(None)
        return df

    def has(self, key):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return key in h5
        return os.path.exists(_key_path(self.path, key, ".table.npz")) or \
            os.path.exists(_key_path(self.path, key, ".npy")) or \
            os.path.exists(_key_path(self.path, key + '/__elements__', ".npz"))

    # ---- plain arrays
    def write_array(self, key, arr, dtype=None):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "a") as h5:
                if key in h5:
                    del h5[key]
                h5.create_dataset(key, data=arr, dtype=dtype)
            return
        np.save(_key_path(self.path, key, ".npy"), np.asarray(arr, dtype=dtype))

    def read_array(self, key):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return h5[key][:]
        return np.load(_key_path(self.path, key, ".npy"), allow_pickle=False)

    def keys(self, prefix=""):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                grp = h5[prefix] if prefix else h5
                return list(grp.keys())
        pre = prefix.strip("/").replace("/", "__")
        out = set()
        for f in os.listdir(self.path):
            base = f[:-len(".table.npz")] if f.endswith(".table.npz") else (f[:-4] if f.endswith(".npy") else None)
            if base is None:
                continue
            if pre:
                if not base.startswith(pre + "__"):
                    continue
                base = base[len(pre) + 2:]
            out.add(base.split("__")[0])
        return sorted(out)

    # ---- per-element groups (window_{W}/<key>/<elt>/{L_counts, region_counts} + attr 'overlaps' in the reference)
    def write_element_groups(self, prefix, names, L_counts, region_counts, overlaps):
        """HDF5: one group per element exactly as preprocess_nonc / preprocess_sites write them
        (sequence_tools.py:639-641, :704-706).  Directory store: one .npz per prefix (a file per element would mean
        hundreds of thousands of files), overlaps as CSR rows of (chrom, start, end)."""
        names = [str(n) for n in names]
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "a") as h5:
                for n, L, R, ov in zip(names, L_counts, region_counts, overlaps):
                    g = '{}/{}'.format(prefix, n)
                    if g in h5:
                        del h5[g]
                    h5.create_dataset(g + '/L_counts', data=L)
                    h5.create_dataset(g + '/region_counts', data=R)
                    h5[g].attrs.create('overlaps', np.asarray(ov))
            return
        ptr = np.zeros(len(names) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(o) for o in overlaps])
        flat = np.array([tuple(int(v) for v in w) for o in overlaps for w in o], dtype=np.int64).reshape(-1, 3)
        np.savez(_key_path(self.path, prefix + '/__elements__', ".npz"), names=np.array(names, dtype=str),
                 L_counts=np.asarray(L_counts, dtype=np.float64), region_counts=np.asarray(region_counts, dtype=np.int64),
                 overlaps_ptr=ptr, overlaps=flat)

    def read_element_groups(self, prefix, names=None):
        """(names, L_counts [E,192], region_counts [E,192], overlaps list) written by write_element_groups; ``names``
        restricts and orders the rows (KeyError for an unknown element, like h5py)."""
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                grp = h5[prefix]
                names = [str(n) for n in (names if names is not None else grp.keys())]
                L = np.array([grp[n]['L_counts'][:] for n in names], dtype=np.float64).reshape(len(names), -1)
                R = np.array([grp[n]['region_counts'][:] for n in names]).reshape(len(names), -1)
                ov = [[tuple(int(v) for v in w) for w in grp[n].attrs['overlaps']] for n in names]
            return names, L, R, ov
        z = np.load(_key_path(self.path, prefix + '/__elements__', ".npz"), allow_pickle=False)
        all_names = [str(n) for n in z["names"]]
        if names is None:
            rows = np.arange(len(all_names))
            names = all_names
        else:
            pos = {n: i for i, n in enumerate(all_names)}
            names = [str(n) for n in names]
            rows = np.array([pos[n] for n in names], dtype=np.int64)
        ptr, flat = z["overlaps_ptr"], z["overlaps"]
        ov = [[tuple(int(v) for v in w) for w in flat[ptr[r]:ptr[r + 1]]] for r in rows]
        return names, z["L_counts"][rows], z["region_counts"][rows], ov

    # ---- attributes
    def _attr_file(self):
        return os.path.join(self.path, "attrs.json")

    def get_attrs(self):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return {k: (v.item() if hasattr(v, "item") else v) for k, v in h5.attrs.items()}
        if os.path.exists(self._attr_file()):
            return json.load(open(self._attr_file()))
        return {}

    def set_attrs(self, **kw):
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "a") as h5:
                for k, v in kw.items():
                    h5.attrs[k] = v
            return
        a = self.get_attrs()
        a.update({k: (v.item() if hasattr(v, "item") else v) for k, v in kw.items()})
        json.dump(a, open(self._attr_file(), "w"), indent=1)


def read_hdf(path, key):
    """Drop-in for the reference's ``pd.read_hdf(path, key)`` call sites."""
    return Store(path, "r").read_table(key)
