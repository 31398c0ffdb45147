"""Collect per-tensor parity errors: one assertion at the end, full table in the failure message and
(on the GPU box) appended to gpurun_out/parity_report.jsonl as evidence."""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Parity:
    def __init__(self, test: str):
        self.test, self.rows = test, []

    def add(self, name: str, err_inf: float, tol: float, err_l2: float = float("nan"), note: str = "") -> None:
        self.rows.append({"test": self.test, "tensor": name, "rel_inf": err_inf, "rel_l2": err_l2, "tol": tol,
                          "ok": bool(err_inf < tol), "note": note})

    def finish(self) -> None:
        out = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "parity_report.jsonl"), "a") as fh:
                for r in self.rows:
                    fh.write(json.dumps(r) + "\n")
        bad = [r for r in self.rows if not r["ok"]]
        table = "\n".join(f"  {'FAIL' if not r['ok'] else 'ok  '} {r['tensor']:<70s} inf={r['rel_inf']:.3e} "
                          f"l2={r['rel_l2']:.3e} tol={r['tol']:.1e} {r['note']}" for r in self.rows)
        assert not bad, f"{self.test}: {len(bad)} tensor(s) out of tolerance\n{table}"
