"""Load tests/golden/golden.json (reference-generated, see tests/golden/make_golden.py)."""
import base64
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
_CACHE = {}


def _synth(spec):
    from pyfastx_b200 import synth
    _, kind, n, seed = spec.split(":")
    n, seed = int(n), int(seed)
    if kind == "fasta":
        return synth.synth_fasta(n, seed=seed)
    if kind == "fasta_w60crlf":
        return synth.synth_fasta(n, seed=seed, min_len=1, max_len=700, width=60, crlf=True)
    if kind == "fastq":
        return synth.synth_fastq(n, seed=seed)
    raise ValueError(spec)


def case_data(case):
    if "data_b64" in case:
        return base64.b64decode(case["data_b64"])
    df = case["datafile"]
    if df.startswith("synth:"):
        return _synth(df)
    return gzip.open(os.path.join(GOLD, "data", df), "rb").read()


def cases(kind=None):
    if "all" not in _CACHE:
        with open(os.path.join(GOLD, "golden.json")) as f:
            _CACHE["all"] = json.load(f)["cases"]
    return [c for c in _CACHE["all"] if kind is None or c["kind"] == kind]


def case_ids(kind=None):
    return [c["name"] for c in cases(kind)]
