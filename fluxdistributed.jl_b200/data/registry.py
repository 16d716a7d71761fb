"""Dataset registry (reference C27 ``Data.toml:1-26`` + DataSets.jl ``BlobTree``).

``dataset("imagenet_local")`` looks a name up in a TOML registry and returns a
*tree*: an object with ``tree[path] -> bytes`` / ``tree.open(path)`` /
``tree.exists(path)``.  Drivers:

* ``FileSystem``  — a directory on disk (``Data.toml:4-12``)
* ``Synthetic``   — procedurally generated ImageNet-shaped content (index CSV,
  synset mapping and JPEG blobs are all fabricated on demand); this is what the
  benchmark configs use because the build has no network and no ImageNet.
* ``S3`` / ``URL`` remote stores (``Data.toml:14-26``): blobs are fetched on demand over HTTP(S) (``file://`` works
  too, which is how the tests exercise the driver without a network) into a local cache directory and served from
  there afterwards.  An ``S3`` entry maps to ``https://<bucket>.s3.amazonaws.com/<prefix>/`` unless ``endpoint`` is
  given.  Without network access the first fetch raises a clear error; cached blobs keep working.

The registry path comes from ``FLUXDIST_DATA_TOML`` or ``Data.toml`` in the CWD /
repo root, and can be configured in code with :func:`register` — the reference's
README promises that API (``README.md:11``).
"""
from __future__ import annotations

import contextlib
import io
import os
import tomllib
import zlib
from typing import Dict, Iterator, Optional

_REGISTRY: Dict[str, dict] = {}
_LOADED_FILES: set = set()


def _repo_root() -> str:
    return os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_registry(path: Optional[str] = None) -> Dict[str, dict]:
    cands = [path] if path else [os.environ.get("FLUXDIST_DATA_TOML"), os.path.join(os.getcwd(), "Data.toml"),
                                 os.path.join(_repo_root(), "Data.toml")]
    for c in cands:
        if c and os.path.exists(c) and c not in _LOADED_FILES:
            with open(c, "rb") as f:
                cfg = tomllib.load(f)
            for ds in cfg.get("datasets", []):
                _REGISTRY.setdefault(ds["name"], ds)
            _LOADED_FILES.add(c)
    return _REGISTRY


def register(name: str, driver: str = "FileSystem", path: Optional[str] = None, **params) -> None:
    """Configure a dataset in code."""
    _REGISTRY[name] = {"name": name, "storage": {"driver": driver, "type": "BlobTree", "path": path, **params}}


class BlobTree:
    """Directory-backed tree."""

    def __init__(self, root: str):
        self.root = root

    def _p(self, path: str) -> str:
        return os.path.join(self.root, path)

    def exists(self, path: str) -> bool:
        return os.path.exists(self._p(path))

    def open(self, path: str, mode: str = "rb"):
        return open(self._p(path), mode)

    def __getitem__(self, path: str) -> bytes:
        with self.open(path) as f:
            return f.read()

    def listdir(self, path: str = ""):
        return sorted(os.listdir(self._p(path)))

    def __repr__(self):
        return f"BlobTree({self.root!r})"


class RemoteTree:
    """URL-backed tree with a read-through local cache (reference: DataSets.jl S3 storage driver, ``Data.toml:14-26``).

    ``tree[path]`` downloads ``<base_url>/<path>`` once (atomically, into ``cache_dir``) and reads the cached copy
    afterwards; ``exists`` checks the cache first, then issues a HEAD / open.  ``timeout`` bounds every request."""

    def __init__(self, base_url: str, cache_dir: Optional[str] = None, timeout: float = 30.0):
        self.base_url = base_url.rstrip("/")
        self.cache_dir = cache_dir or os.path.join(os.path.expanduser("~"), ".cache", "fluxdist", "datasets",
                                                   format(zlib.crc32(self.base_url.encode()), "08x"))
        self.timeout = timeout
        os.makedirs(self.cache_dir, exist_ok=True)

    def _local(self, path: str) -> str:
        return os.path.join(self.cache_dir, path.lstrip("/"))

    def _fetch(self, path: str) -> str:
        import tempfile
        import urllib.error
        import urllib.request

        loc = self._local(path)
        if os.path.exists(loc):
            return loc
        url = f"{self.base_url}/{path.lstrip('/')}"
        os.makedirs(os.path.dirname(loc) or ".", exist_ok=True)
        try:
            with urllib.request.urlopen(url, timeout=self.timeout) as r:
                data = r.read()
        except (urllib.error.URLError, OSError, ValueError) as e:
            raise FileNotFoundError(f"remote dataset blob {url!r} could not be fetched ({e}); no cached copy at {loc!r}") from e
        fd, tmp = tempfile.mkstemp(dir=os.path.dirname(loc) or ".", prefix=".blob-")
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        os.replace(tmp, loc)  # atomic: concurrent loader threads never see a partial blob
        return loc

    def exists(self, path: str) -> bool:
        try:
            self._fetch(path)
            return True
        except FileNotFoundError:
            return False

    def open(self, path: str, mode: str = "rb"):
        return open(self._fetch(path), mode)

    def __getitem__(self, path: str) -> bytes:
        with self.open(path) as f:
            return f.read()

    def __repr__(self):
        return f"RemoteTree({self.base_url!r}, cache={self.cache_dir!r})"


class SyntheticTree:
    """ImageNet-shaped tree generated on the fly.

    * ``LOC_synset_mapping.txt`` — ``nclasses`` fake synsets
    * ``LOC_train_solution.csv`` / ``LOC_val_solution.csv`` — ``ntrain`` / ``nval`` rows
    * ``ILSVRC/Data/CLS-LOC/{train/<wnid>/,val/}<id>.JPEG`` — a deterministic
      pseudo-random image per id (JPEG-encoded when PIL is present so the decode
      path is exercised; otherwise a raw ``.npy`` payload).
    """

    def __init__(self, nclasses: int = 1000, ntrain: int = 10000, nval: int = 1000, size=(256, 320),
                 encode_jpeg: bool = True):
        self.nclasses, self.ntrain, self.nval = nclasses, ntrain, nval
        self.size = size
        self.encode_jpeg = encode_jpeg

    def wnid(self, c: int) -> str:
        return f"n{c + 1:08d}"

    def _solution(self, split: str, n: int) -> str:
        rows = ["ImageId,PredictionString"]
        for i in range(n):
            c = (i * 7919) % self.nclasses
            w = self.wnid(c)
            iid = f"{w}_{i}" if split == "train" else f"ILSVRC2012_val_{i + 1:08d}"
            rows.append(f"{iid},{w} 10 10 200 200")
        return "\n".join(rows) + "\n"

    def _image(self, path: str) -> bytes:
        import numpy as np

        seed = zlib.crc32(path.encode())
        rng = np.random.default_rng(seed)
        h, w = self.size
        base = rng.integers(0, 255, size=(h // 16 + 1, w // 16 + 1, 3), dtype=np.uint8)
        img = np.kron(base, np.ones((16, 16, 1), dtype=np.uint8))[:h, :w]
        if self.encode_jpeg:
            try:
                from PIL import Image

                buf = io.BytesIO()
                Image.fromarray(img).save(buf, format="JPEG", quality=85)
                return buf.getvalue()
            except Exception:
                pass
        buf = io.BytesIO()
        np.save(buf, img)
        return buf.getvalue()

    def exists(self, path: str) -> bool:
        return True

    def __getitem__(self, path: str) -> bytes:
        if path.endswith("LOC_synset_mapping.txt"):
            return "".join(f"{self.wnid(c)} synthetic class {c}, fake {c}\n" for c in range(self.nclasses)).encode()
        if path.endswith("LOC_train_solution.csv"):
            return self._solution("train", self.ntrain).encode()
        if path.endswith("LOC_val_solution.csv"):
            return self._solution("val", self.nval).encode()
        return self._image(path)

    def open(self, path: str, mode: str = "rb"):
        data = self[path]
        return io.BytesIO(data) if "b" in mode else io.StringIO(data.decode())

    def __repr__(self):
        return f"SyntheticTree(nclasses={self.nclasses}, ntrain={self.ntrain})"


def dataset(name: str):
    """``DataSets.dataset(name)`` -> an openable tree."""
    load_registry()
    if name not in _REGISTRY:
        if name.startswith("synthetic"):
            return SyntheticTree()
        raise KeyError(f"dataset {name!r} not in registry ({sorted(_REGISTRY)}); edit Data.toml or call "
                       f"fluxdist.data.register(name, path=...)")
    st = _REGISTRY[name].get("storage", {})
    drv = st.get("driver", "FileSystem")
    if drv == "FileSystem":
        path = os.path.expanduser(st.get("path") or "")
        if not os.path.isdir(path):
            raise FileNotFoundError(f"dataset {name!r}: directory {path!r} does not exist (edit Data.toml)")
        return BlobTree(path)
    if drv == "Synthetic":
        kw = {k: st[k] for k in ("nclasses", "ntrain", "nval", "encode_jpeg") if k in st}
        if "size" in st:
            kw["size"] = tuple(st["size"])
        return SyntheticTree(**kw)
    if drv in ("S3", "URL", "HTTP"):
        base = st.get("url") or st.get("endpoint")
        if base is None and drv == "S3":
            base = f"https://{st['bucket']}.s3.amazonaws.com"
        if base is None:
            raise ValueError(f"dataset {name!r}: the {drv} driver needs `url` (or `bucket` for S3)")
        if st.get("prefix"):
            base = base.rstrip("/") + "/" + str(st["prefix"]).strip("/")
        return RemoteTree(base, cache_dir=os.path.expanduser(st["cache"]) if st.get("cache") else None,
                          timeout=float(st.get("timeout", 30.0)))
    raise RuntimeError(f"dataset {name!r}: unknown storage driver {drv!r}")


@contextlib.contextmanager
def open_tree(ds) -> Iterator:
    """``open(BlobTree, dataset(name)) do tree ... end``."""
    yield dataset(ds) if isinstance(ds, str) else ds
