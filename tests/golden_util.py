"""Access to the committed reference outputs under tests/golden/ (made by tests/golden/make_golden.py)."""
import gzip
import json
import os
import shutil

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["quirk", "dense", "cfg1"]


class GoldenCase:
    def __init__(self, name: str, workdir: str):
        self.name = name
        src = os.path.join(GOLDEN, name)
        self.dir = os.path.join(workdir, "golden_" + name)
        if not os.path.isdir(self.dir):
            os.makedirs(self.dir)
            with gzip.open(os.path.join(src, "db.igd.gz"), "rb") as g, open(os.path.join(self.dir, "db.igd"), "wb") as f:
                shutil.copyfileobj(g, f)
            for fn in ("db_index.tsv", "q.bed", "qf.bed"):
                if os.path.exists(os.path.join(src, fn)):
                    shutil.copy(os.path.join(src, fn), os.path.join(self.dir, fn))
        self.igd = os.path.join(self.dir, "db.igd")
        self.qfile = os.path.join(self.dir, "q.bed")
        self.exp = json.load(open(os.path.join(src, "expected.json")))
        self.src = src

    def hits(self, v: int = 0) -> np.ndarray:
        return np.array(self.exp["hits" if v == 0 else f"hits_v{v}"], np.int64)

    def ret(self, v: int = 0) -> int:
        return self.exp["ret" if v == 0 else f"ret_v{v}"]

    def seqpare(self) -> np.ndarray:
        return np.array([float.fromhex(x) for x in self.exp["seqpare_hex"]], np.float64)

    def hitsmap(self, v: int = 0) -> np.ndarray:
        return np.array(self.exp["hitsmap" if v == 0 else f"hitsmap_v{v}"], np.uint32)

    def stdout(self, key: str) -> str:
        with gzip.open(os.path.join(self.src, f"stdout_{key}.txt.gz"), "rt") as f:
            return f.read()

    def cli_args(self, key: str) -> list:
        """argv after `igd search db.igd`, with the query files made absolute"""
        # (cases without a reduced -f query file were recorded with "qf.bed" standing for q.bed itself)
        qf = os.path.join(self.dir, "qf.bed")
        m = {"q.bed": self.qfile, "qf.bed": qf if os.path.exists(qf) else self.qfile}
        return [m.get(a, a) for a in self.exp["cli_args"][key]]

    def normalise(self, text: str) -> str:
        return text.replace(self.qfile, "q.bed").replace(os.path.join(self.dir, "qf.bed"), "qf.bed")
