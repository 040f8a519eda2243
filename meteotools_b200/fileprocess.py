"""load_settings of the reference (meteotools/fileprocess.py:37-40); file/time helpers that only
serve wrfout file discovery are out of scope."""
import json


def load_settings(settings_path):
    with open(settings_path, 'r') as f:
        return json.load(f)
