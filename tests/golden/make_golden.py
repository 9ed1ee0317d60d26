"""Regenerates tests/golden/ from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, which is why the outputs are committed).

  python tests/golden/make_golden.py

* inputs/*.ppm       : the reference's small P3 fixtures, verbatim (tests/small.ppm, 8x8.ppm,
                       16x16.ppm [really 8x8], 7x17.ppm)
* inputs/500x500.png : tests/500x500.ppm pixels, losslessly re-packed (the P3 text is 2.25 MB)
* jpeg/<name>_<preset>.jpg : the oracle's output for every fixture x preset
* expected_sha256.json : SHA-256 of those files, as listed in SURVEY.md section 8c.  That table
                       was produced by an INDEPENDENT numpy restatement during the survey;
                       this script asserts the C oracle reproduces it bit for bit.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/tests"
SURVEY_SHA = {
    "small_P420": "dd798df92163a5c58ff41a3d1c6bbf1d8b1bc89062cd680cfbf1a3fa2cc49ecf",
    "small_P422": "eb51fed551f3a27890062441c2c5ee9a75a91bd07109e62ffb1060f1038ef422",
    "small_P444": "9d591f67fef7aee9db44d0c2cb83db6e9ab2eeee6d475b2ed6606f368eab34a5",
    "8x8_P420": "1dbfceeae2a0a9a708e8855989227c57bc40907d742ee1a64c3c0988feee0438",
    "8x8_P422": "471b0ded5559a72d839bf229a5bf3a178b4c4a130ab61fb3c24dc14d247a1cd9",
    "8x8_P444": "b9ffdb7cf89815d4ce10958dd427ff0eb6d970f05a26459e5f997ecbc08c6db6",
    "16x16_P420": "1c25c583e1427ca8adef85024fefe12f100cb2c5795562ce0b1e65af3ccb3fa9",
    "16x16_P422": "9e67032bd2304112cafd00ce50b3948106c0b7893001827b79bc08a59a5db352",
    "16x16_P444": "5dbe357e3dd422600a93feeb421eeb485c5a9674d71ad433fe780b5f32c43593",
    "7x17_P420": "4b8653fc442b334345c2bbcf8a890e5808631201d255f335708f883d1b54b136",
    "7x17_P422": "ab16c4070a5b04e90c5152759ae0c1c94e617c64387f79f5790a30feefdbd5df",
    "7x17_P444": "38c7ecc149211a20e41da8f0658aadac336d63acd5c6754d6c202bf1b142e313",
    "500x500_P420": "ddc048221477660309cc773ca3171576080be02a9def3c41a97cb1f62925ee28",
    "500x500_P422": "235fa5d51c0f9626a9b0b70896d8faa6b96a14e54622f34c70e84994abad0504",
    "500x500_P444": "1682183552533f5f2bf7e3aa45576abd11e0126687680c1e38f8f7cb90431140",
}
PRESETS = {"P444": O.P444, "P422": O.P422, "P420": O.P420}


def main():
    for name in ("small", "8x8", "16x16", "7x17"):
        shutil.copyfile(f"{REF}/{name}.ppm", f"{HERE}/inputs/{name}.ppm")
    w, h, mx, px = O.parse_ppm(open(f"{REF}/500x500.ppm", "rb").read())
    assert (w, h, mx) == (500, 500, 255)
    Image.fromarray(px.astype(np.uint8), "RGB").save(f"{HERE}/inputs/500x500.png", optimize=True)
    for name in ("small", "8x8", "16x16", "7x17", "500x500"):
        text = open(f"{REF}/{name}.ppm", "rb").read()
        for pname, p in PRESETS.items():
            r = O.encode_ppm(text, preset=p)
            key = f"{name}_{pname}"
            got = hashlib.sha256(r.jpeg).hexdigest()
            assert got == SURVEY_SHA[key], (key, got)
            open(f"{HERE}/jpeg/{key}.jpg", "wb").write(r.jpeg)
    json.dump(SURVEY_SHA, open(f"{HERE}/expected_sha256.json", "w"), indent=1, sort_keys=True)
    print("golden regenerated; all", len(SURVEY_SHA), "hashes match the survey table")


if __name__ == "__main__":
    main()
