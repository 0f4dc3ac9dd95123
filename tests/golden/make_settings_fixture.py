"""Generates tests/golden/reference_settings.json: the key/value content of every settings/*.json the reference ships
(/root/reference/settings/), as input vectors for the settings-surface test (tests/test_host_logic.py).  Run in the
build container (the reference tree does not exist on the GPU box):  python tests/golden/make_settings_fixture.py"""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
out = {os.path.basename(f): json.load(open(f)) for f in sorted(glob.glob("/root/reference/settings/*.json"))}
with open(os.path.join(HERE, "reference_settings.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(len(out), "settings files")
