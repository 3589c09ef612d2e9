"""configs/*.yaml loader: the reference reads its YAML through OmegaConf and uses attribute access
(train.py:223-226, model/titok.py:28-42). OmegaConf is optional here: a plain attribute dict over yaml.safe_load
accepts the same files verbatim."""
from __future__ import annotations

import yaml


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(o):
        if isinstance(o, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in o.items()})
        if isinstance(o, list):
            return [AttrDict.wrap(v) for v in o]
        return o


def load_config(path: str) -> AttrDict:
    with open(path) as f:
        return AttrDict.wrap(yaml.safe_load(f))


def tiny_config(fsq_levels=(7, 5, 5, 5, 5), patch_size=(4, 8, 8), encoder_size="tiny", decoder_size="tiny") -> AttrDict:
    """The `tokenizer.model` block of the reference's configs/tiny.yaml:14-19 (the only keys the model reads)."""
    return AttrDict.wrap({"tokenizer": {"model": {"patch_size": list(patch_size), "fsq_levels": list(fsq_levels),
                                                  "encoder_size": encoder_size, "decoder_size": decoder_size}}})
