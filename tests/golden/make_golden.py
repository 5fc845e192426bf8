"""Regenerates tests/golden/splitmix64.json and sanity-checks gob_doc_vectors.json (hand-transcribed
from the encoding/gob package documentation; the reference cannot be run here)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
M = (1 << 64) - 1


def splitmix64(seed, i):
    z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def main():
    out = {"algorithm": "splitmix64 (Steele, Lea, Flood 2014): state += 0x9E3779B97F4A7C15 per output",
           "seed0": ["%016x" % splitmix64(0, i) for i in range(8)],
           "seed_b2000000": ["%016x" % splitmix64(0xB2000000, i) for i in range(8)]}
    with open(os.path.join(HERE, "splitmix64.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(HERE, "gob_doc_vectors.json")) as f:
        doc = json.load(f)
    td = bytes.fromhex(doc["point_type_descriptor"])
    assert td[0] == len(td) - 1 == 0x1F, "descriptor message is 31 bytes long"
    val = bytes.fromhex(doc["point_value_22_33"])
    assert val[0] == len(val) - 1 == 7 and val[4] >> 1 == 22 and val[6] >> 1 == 33
    print("ok")


if __name__ == "__main__":
    main()
