"""Small helpers matching /root/reference/src/util.py:23-39."""
import json
import logging

import numpy as np


def load_experiment_parameters(parameters_path):
    try:
        with open(parameters_path, "r") as fin:
            return json.load(fin)
    except FileNotFoundError:
        logging.warning("File '%s' not found.", parameters_path)
        return {}


def normalize(arr):
    if arr.ndim == 1:
        return arr / np.linalg.norm(arr)
    return arr / np.linalg.norm(arr, axis=1, keepdims=True)
