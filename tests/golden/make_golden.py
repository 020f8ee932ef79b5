#!/usr/bin/env python
"""Pack the reference's committed golden vectors into compact fixtures.

Run in the build container only (it reads ``/root/reference``, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

Reads the librosa-0.11 / SoXR golden JSON files of the reference test suites

    soundml/test/window/vectors/*.json
    soundml/test/stft/vectors/*.json
    soundml/test/mel/vectors/{filterbank,mel_spectrogram,mfcc}.json
    soundml/test/db/vectors/*.json
    soundml/test/istft/vectors/{inverse_*,lengths,griffinlim_*}.json
    soundml/test/resample/vectors/soxr_reference.json

and writes ``tests/golden/reference_vectors.npz``: one float64 array per case
(key ``<suite>/<file>/<case>``) plus a JSON index (key ``__index__``) holding
each case's parameters and shape.  No reference *source* is copied, only the
numeric test data the reference's own tests replay.
"""
import glob
import json
import os
import sys

import numpy as np

REF = "/root/reference/soundml/test"
HERE = os.path.dirname(os.path.abspath(__file__))

SUITES = {
    "window": sorted(glob.glob(f"{REF}/window/vectors/*.json")),
    "stft": sorted(glob.glob(f"{REF}/stft/vectors/*.json")),
    "mel": [f"{REF}/mel/vectors/filterbank.json",
            f"{REF}/mel/vectors/mel_spectrogram.json",
            f"{REF}/mel/vectors/mfcc.json"],
    "db": sorted(glob.glob(f"{REF}/db/vectors/*.json")),
    "istft": sorted(glob.glob(f"{REF}/istft/vectors/inverse_*.json")) +
             [f"{REF}/istft/vectors/lengths.json"],
    "griffinlim": sorted(glob.glob(f"{REF}/istft/vectors/griffinlim_*.json")),
    "resample": [f"{REF}/resample/vectors/soxr_reference.json"],
}


def main():
    arrays, index = {}, {}
    for suite, files in SUITES.items():
        for path in files:
            stem = os.path.splitext(os.path.basename(path))[0]
            doc = json.load(open(path))
            for case in doc["cases"]:
                key = f"{suite}/{stem}/{case['name']}"
                params = dict(case["params"])
                if "input" in params:          # db suite: the input rides in the params
                    arrays[key + "#input"] = np.asarray(params.pop("input"), dtype=np.float64)
                entry = {"params": params}
                if "values" in case:
                    arrays[key] = np.asarray(case["values"], dtype=np.float64)
                    entry["shape"] = case["shape"]
                else:                      # window/cola.json: boolean verdicts
                    entry["expected"] = case["expected"]
                index[key] = entry
    arrays["__index__"] = np.frombuffer(
        json.dumps(index, sort_keys=True).encode(), dtype=np.uint8)
    out = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(out, **arrays)
    print(f"wrote {out}: {len(index)} cases, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not mounted; fixtures are already committed")
    main()
