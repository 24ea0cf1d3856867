"""Reader for the velocity maps the CLI writes: the eight vmap_<t0>_<t1>_<field>.GMA files plus
_meta.txt, loose in a directory or packed in the .tar the driver makes (MIMC_main.c:428-480).
Same fields and the same control-point bias removal as the reference's vmap.py:10-185, without
its plotting / MATLAB export."""
from __future__ import annotations

import glob
import os
import tarfile

import numpy as np

from . import gma

FIELDS = ("x", "y", "vx", "vy", "ex", "ey", "qual", "flagcp")


def parse_meta(text: str) -> dict:
    """key=value lines; cp_offset_* are numbers (int when they parse as int), the rest strings."""
    meta = {}
    for line in text.splitlines():
        if not line.strip() or "=" not in line:
            continue
        k, v = line.split("=", 1)
        if "cp_offset" in k:
            try:
                meta[k] = int(v)
            except ValueError:
                meta[k] = float(v)
        else:
            meta[k] = v
    return meta


class VMap:
    def __init__(self, path: str):
        """path: output directory of a run, or the vmap_*.tar."""
        self.path = path
        self._data = {}
        self._tar = path if os.path.isfile(path) and tarfile.is_tarfile(path) else None

    def _raw(self, suffix: str) -> bytes:
        if self._tar:
            with tarfile.open(self._tar) as tf:
                names = [m for m in tf.getmembers() if m.name.endswith(suffix)]
                if len(names) != 1:
                    raise FileNotFoundError(f"{self._tar}: expected one member ending in {suffix}, found {len(names)}")
                return tf.extractfile(names[0]).read()
        hits = glob.glob(os.path.join(self.path, "*" + suffix))
        if len(hits) != 1:
            raise FileNotFoundError(f"{self.path}: expected one file ending in {suffix}, found {len(hits)}")
        with open(hits[0], "rb") as f:
            return f.read()

    def field(self, name: str) -> np.ndarray:
        if name not in self._data:
            if name not in FIELDS:
                raise KeyError(name)
            self._data[name] = gma.read(self._raw(f"_{name}.GMA"), gma.FIELD_DTYPE[name])
        return self._data[name]

    x = property(lambda s: s.field("x")); y = property(lambda s: s.field("y"))
    vx = property(lambda s: s.field("vx")); vy = property(lambda s: s.field("vy"))
    ex = property(lambda s: s.field("ex")); ey = property(lambda s: s.field("ey"))
    qual = property(lambda s: s.field("qual")); flagcp = property(lambda s: s.field("flagcp"))

    @property
    def spd(self) -> np.ndarray:
        return np.sqrt(self.vx ** 2 + self.vy ** 2)

    @property
    def meta(self) -> dict:
        if "meta" not in self._data:
            self._data["meta"] = parse_meta(self._raw("_meta.txt").decode())
        return self._data["meta"]

    def adjust(self):
        """Subtract the mean velocity over the control points (vmap.py:174-185) -> (bias_vx, bias_vy)."""
        cp = self.flagcp != 0
        with np.errstate(invalid="ignore"):
            bx = float(np.nanmean(self.vx[cp])) if cp.any() else 0.0
            by = float(np.nanmean(self.vy[cp])) if cp.any() else 0.0
        self._data["vx"] = self.vx - np.float32(bx)
        self._data["vy"] = self.vy - np.float32(by)
        return bx, by
