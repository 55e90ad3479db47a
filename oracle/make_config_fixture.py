"""Dump the reference's JSON configs (key -> value per file) into tests/golden/config_contract.json so that the
config-contract test can run where /root/reference is absent.  Build-container only; test infrastructure."""
import glob
import json
import os

REF = "/root/reference/dmi/configs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "config_contract.json")
out = {}
for path in sorted(glob.glob(os.path.join(REF, "**", "*.json"), recursive=True)):
    rel = os.path.relpath(path, REF)
    txt = open(path).read().strip()
    if not txt:
        continue                      # dmi/configs/config.json is empty in the reference
    out[rel] = json.loads(txt)
json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
print(len(out), "configs ->", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")
