"""JSON config -> nested attribute object; same schema and behaviour as the reference's
trainer/config.py:10-22 (``load_JsonConfig(path).Model.vq_path`` etc.), so the reference's own
config/*.json files load unchanged."""
import json


class Object:
    def __init__(self, config: dict) -> None:
        for key, value in config.items():
            setattr(self, key, Object(value) if isinstance(value, dict) else value)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Object) else v) for k, v in self.__dict__.items()}


def load_JsonConfig(json_file):
    with open(json_file, "r") as f:
        return Object(json.load(f))
