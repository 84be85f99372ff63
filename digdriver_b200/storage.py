"""Interchange files of the hot path.

The reference hands data between its CLI stages through HDF5 files written with h5py and
``pandas.to_hdf`` (region_params, sequence_model_192/64, genome_counts, all_window_genome_counts, idx,
mappability, window_{W}/..., pretrained tables; DigPreprocess.py:63-73, DigPretrain.py:82-96,156-177,
207-208).  Neither h5py nor PyTables exists in this image, so the same key layout is served by a
directory store (``<path>`` is a directory holding ``<key>.npy`` / ``<key>.table.npz`` files and
``attrs.json``).  An existing HDF5 file is read through h5py + PyTables when they are importable and otherwise
through ``hdf5_lite`` (a pure-Python reader of the classic HDF5 structures and the pandas fixed format; read-only),
so pretrained models produced by the reference can be dropped in.  ``export_hdf5`` writes a directory store out as
an HDF5 file with the reference's key layout.
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


def _owned_file(name):
    """Files a directory store writes: <key>.npy, <key>.table.npz, <prefix>____elements__.npz, attrs.json."""
    return name == "attrs.json" or name.endswith(".npy") or name.endswith(".table.npz") or \
        name.endswith("____elements__.npz")


class Store:
    """Directory-backed key/value store with the reference's HDF5 key names."""

    def __init__(self, path, mode="a"):
        self.path = str(path)
        self.hdf5 = _is_hdf5_file(self.path)
        self.lite = None
        if self.hdf5 and not _have_hdf5():
            if mode != "r":
                raise RuntimeError("%s is an HDF5 file and h5py/PyTables are not installed: it can only be opened "
                                   "read-only (mode 'r') through the built-in reader" % self.path)
            from . import hdf5_lite
            self.lite = hdf5_lite.File(self.path)
            self.hdf5 = False
        if not self.hdf5 and self.lite is None:
            if mode == "r" and not os.path.isdir(self.path):
                raise FileNotFoundError(self.path)
            if mode == "w" and os.path.isdir(self.path):
                # truncate like HDF5 mode 'w' would -- but only files this store owns: never a user's directory
                entries = os.listdir(self.path)
                foreign = [f for f in entries if not _owned_file(f)]
                if foreign:
                    raise RuntimeError("%s exists and holds files that are not part of a store (%s%s): refusing to "
                                       "truncate it; choose another output path" %
                                       (self.path, ", ".join(sorted(foreign)[:3]), ", ..." if len(foreign) > 3 else ""))
                for f in entries:
                    os.remove(os.path.join(self.path, f))
            os.makedirs(self.path, exist_ok=True)

    # ---- tables (pandas objects)
    def _writable(self):
        if self.lite is not None:
            raise RuntimeError("%s was opened through the built-in read-only HDF5 reader" % self.path)

    def write_table(self, key, df):
        self._writable()
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
        # tuple column labels (the (MUT_TYPE, CONTEXT) columns of the reference's mutation-count tables) survive as JSON
        tuples = json.dumps([list(map(str, c)) for c in df.columns]) if any(isinstance(c, tuple) for c in df.columns) else ""
        np.savez(_key_path(self.path, key, ".table.npz"), __columns__=np.array([str(c) for c in df.columns]),
                 __index__=plain(df.index), __index_name__=np.array([str(df.index.name or "")]),
                 __column_tuples__=np.array([tuples]), **cols)

    def read_table(self, key):
        if self.lite is not None:
            return self.lite.read_pandas(key)
        if self.hdf5:
            return pd.read_hdf(self.path, key)
        z = np.load(_key_path(self.path, key, ".table.npz"), allow_pickle=False)
        cols = [str(c) for c in z["__columns__"]]
        df = pd.DataFrame({c: z["col%d" % i] for i, c in enumerate(cols)}, index=z["__index__"])
        name = str(z["__index_name__"][0])
        df.index.name = name or None
        tuples = str(z["__column_tuples__"][0]) if "__column_tuples__" in z.files else ""
        if tuples:
            df.columns = pd.MultiIndex.from_tuples([tuple(c) for c in json.loads(tuples)])
        if cols == ["__series__"]:
            return df["__series__"].rename(None)
        return df

    def has(self, key):
        if self.lite is not None:
            return key in self.lite
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return key in h5
        return os.path.exists(_key_path(self.path, key, ".table.npz")) or \
            os.path.exists(_key_path(self.path, key, ".npy")) or \
            os.path.exists(_key_path(self.path, key + '/__elements__', ".npz"))

    # ---- plain arrays
    def write_array(self, key, arr, dtype=None):
        self._writable()
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "a") as h5:
                if key in h5:
                    del h5[key]
                h5.create_dataset(key, data=arr, dtype=dtype)
            return
        np.save(_key_path(self.path, key, ".npy"), np.asarray(arr, dtype=dtype))

    def read_array(self, key):
        if self.lite is not None:
            return self.lite.read(key)
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return h5[key][:]
        return np.load(_key_path(self.path, key, ".npy"), allow_pickle=False)

    def keys(self, prefix=""):
        if self.lite is not None:
            return self.lite.keys(prefix or "/")
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
        self._writable()
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
        if self.lite is not None:
            f = self.lite
            names = [str(n) for n in (names if names is not None else f.keys(prefix))]
            L = np.array([f.read('{}/{}/L_counts'.format(prefix, n)) for n in names], dtype=np.float64).reshape(len(names), -1)
            R = np.array([f.read('{}/{}/region_counts'.format(prefix, n)) for n in names]).reshape(len(names), -1)
            ov = [[tuple(int(v) for v in w) for w in np.asarray(f.attrs('{}/{}'.format(prefix, n))['overlaps']).reshape(-1, 3)]
                  for n in names]
            return names, L, R, ov
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
        if self.lite is not None:
            return {k: (v.item() if hasattr(v, "item") and getattr(v, "ndim", 1) == 0 else v)
                    for k, v in self.lite.attrs("/").items()}
        if self.hdf5:
            import h5py
            with h5py.File(self.path, "r") as h5:
                return {k: (v.item() if hasattr(v, "item") else v) for k, v in h5.attrs.items()}
        if os.path.exists(self._attr_file()):
            return json.load(open(self._attr_file()))
        return {}

    def set_attrs(self, **kw):
        self._writable()
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


def _pandas_node(obj):
    """The group pandas' fixed format writes for a DataFrame / Series (pandas/io/pytables.py, BlockManagerFixed /
    SeriesFixed): axis*/block*_items index arrays with a 'kind' attribute, block*_values stored transposed.  String
    (object) blocks are written as fixed-length byte arrays, not as PyTables' pickled VLArrays."""
    def index_ds(values, name=None):
        v = np.asarray(values)
        kind = "string" if v.dtype.kind in "OUS" else ("integer" if v.dtype.kind in "iu" else "float")
        return ('data', v, {"kind": kind, "name": "N." if name is None else str(name), "transposed": np.bool_(True)})
    common = {"CLASS": "GROUP", "TITLE": "", "VERSION": "1.0", "encoding": "UTF-8", "errors": "strict",
              "pandas_version": "0.15.2"}
    if isinstance(obj, pd.Series):
        return {"attrs": dict(common, pandas_type="series", name="N." if obj.name is None else str(obj.name)),
                "children": {"index": index_ds(obj.index.values, obj.index.name),
                             "values": ('data', np.asarray(obj.values), {"transposed": np.bool_(True)})}}
    children = {"axis0": index_ds([str(c) for c in obj.columns], obj.columns.name),
                "axis1": index_ds(obj.index.values, obj.index.name)}
    blocks = {}
    for c in obj.columns:
        v = np.asarray(obj[c].values)
        k = "str" if v.dtype.kind in "OUS" else str(v.dtype)
        blocks.setdefault(k, []).append(c)
    attrs = dict(common, pandas_type="frame", ndim=np.int64(2), nblocks=np.int64(len(blocks)),
                 axis0_variety="regular", axis1_variety="regular")
    for i, (k, cols) in enumerate(blocks.items()):
        vals = np.stack([np.asarray(obj[c].values) if k != "str" else np.array([str(x) for x in obj[c].values])
                         for c in cols], axis=1) if len(obj) else np.zeros((0, len(cols)))
        children["block%d_items" % i] = index_ds([str(c) for c in cols])
        children["block%d_values" % i] = ('data', vals, {"transposed": np.bool_(True)})
        attrs["block%d_items_variety" % i] = "regular"
    return {"attrs": attrs, "children": children}


def export_hdf5(store_path, out_path):
    """Write a directory store as ONE HDF5 file with the reference's key layout (datasets for arrays, pandas
    fixed-format groups for tables, root attributes, per-element groups with L_counts / region_counts / overlaps).
    The file is produced by hdf5_lite's classic-format writer; it has been read back with hdf5_lite only (libhdf5 is
    not available in this image), see hdf5_lite's module docstring."""
    from . import hdf5_lite
    st = Store(store_path, "r")
    root = {"attrs": st.get_attrs(), "children": {}}

    def node_for(parts):
        node = root
        for part in parts:
            node = node["children"].setdefault(part, {"attrs": {}, "children": {}})
        return node
    for f in sorted(os.listdir(st.path)):
        if f.endswith(".table.npz"):
            parts = f[:-len(".table.npz")].split("__")
            node_for(parts[:-1])["children"][parts[-1]] = _pandas_node(st.read_table("/".join(parts)))
        elif f.endswith(".npy"):
            parts = f[:-4].split("__")
            node_for(parts[:-1])["children"][parts[-1]] = ('data', st.read_array("/".join(parts)), {})
        elif f.endswith("____elements__.npz"):
            parts = [x for x in f[:-len("____elements__.npz")].split("__")]
            names, L, R, ov = st.read_element_groups("/".join(parts))
            grp = node_for(parts)
            for n, l_, r_, o_ in zip(names, L, R, ov):
                grp["children"][n] = {"attrs": {"overlaps": np.asarray(o_, dtype=np.int64).reshape(-1, 3)},
                                      "children": {"L_counts": ('data', l_, {}), "region_counts": ('data', r_, {})}}
    hdf5_lite.hdf5_write(out_path, root)
    return out_path
