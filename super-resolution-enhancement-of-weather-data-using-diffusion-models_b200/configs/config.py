"""Experiment configuration -- same schema and CLI semantics as the reference's configs/config.py:49-133
(JSON with ``//`` comments stripped line-wise, experiment directory creation, GPU selection).

Kept quirk: ``params['gpu_ids']`` ends up as the raw ``-gpu`` string (reference config.py:65).  Changed on purpose: the
B200 path has no CPU mode, so omitting ``-gpu`` falls back to the ids in the JSON instead of selecting the CPU, and
multi-GPU means one process per GPU (``torchrun``), not ``nn.DataParallel``.
Optional extra key: ``model.precision`` = ``"bf16"`` (default) | ``"fp32"`` (check mode).
"""
import json
import os
from collections import OrderedDict
from datetime import datetime


def strip_comments(text):
    """Everything after the first ``//`` on a line is a comment (the reference's rule, config.py:82-87)."""
    return "\n".join(line.split("//")[0] for line in text.splitlines())


def load_json_with_comments(path):
    with open(path, "r") as fh:
        return json.loads(strip_comments(fh.read()), object_pairs_hook=OrderedDict)


def get_current_datetime() -> str:
    return datetime.now().strftime("%y%m%d_%H%M%S")


def mkdirs(paths) -> None:
    for p in ([paths] if isinstance(paths, str) else paths):
        os.makedirs(p, exist_ok=True)


class DataConfig:
    """On-disk store conventions (reference configs/config.py:8-46 + configs/data_config/config.json).  The reference reads
    them from a JSON file next to its config module; the same keys can be overridden here with ``json_path``."""

    DEFAULTS = {"name": "data_config", "datetime_format": "%Y-%m-%d-%H", "temporal_resolution": {"unit": "h", "value": 1},
                "directory_name_meta_data": "meta", "file_name_meta_data": "metadata", "file_name_constant_data": "constant",
                "directory_name_sample_data": "samples", "netcdf_extension": ".nc", "numpy_extension": ".npy"}

    def __init__(self, json_path=None, json_name=None):
        cfg = dict(self.DEFAULTS)
        if json_name and not json_path:
            json_path = os.path.join(os.path.dirname(__file__), "data_config", (json_name[:-5] if json_name.endswith(".json") else json_name) + ".json")
            if not os.path.exists(json_path):
                raise FileNotFoundError("Json file with given name does not exist: {}".format(json_path))
        if json_path:
            with open(json_path, "r") as fh:
                cfg.update(json.load(fh))
        self.config = cfg
        self.name = cfg["name"]
        self.datetime_format = cfg["datetime_format"]
        self.temporal_resolution_unit = cfg["temporal_resolution"]["unit"]
        self.temporal_resolution_value = cfg["temporal_resolution"]["value"]
        self.directory_name_meta_data = cfg["directory_name_meta_data"]
        self.file_name_meta_data = cfg["file_name_meta_data"]
        self.file_name_constant_data = cfg["file_name_constant_data"]
        self.directory_name_sample_data = cfg["directory_name_sample_data"]
        self.netcdf_extension = cfg["netcdf_extension"]
        self.numpy_extension = cfg["numpy_extension"]


class Config:
    def __init__(self, args, experiment=True):
        self.args = args
        self.root = args.config
        self.gpu_ids = getattr(args, "gpu_ids", None)
        self.experiments_root = None
        self.params = load_json_with_comments(self.root)
        if experiment:
            self.handle_experiment_configs()
        if self.gpu_ids:
            gpu_list = self.gpu_ids
        else:
            gpu_list = ",".join(str(x) for x in self.params.get("gpu_ids", [0]))
        if "LOCAL_RANK" not in os.environ:            # under torchrun every rank keeps all devices visible
            os.environ["CUDA_VISIBLE_DEVICES"] = gpu_list
        self.params["distributed"] = len(gpu_list) > 1
        self.params["gpu_ids"] = self.gpu_ids if self.gpu_ids else gpu_list
        if "phase" in vars(args) and getattr(args, "phase", None):
            self.params["phase"] = args.phase
        tg = self.params.get("data", {}).get("transform_groups")
        if isinstance(tg, dict):
            self.params["data"]["transform_groups"] = list(tg.values())

    def get_opt(self):
        return self.params

    def get_hyperparameters_as_dict(self):
        return self.params

    def handle_experiment_configs(self):
        path = self.params["path"]
        if not path.get("resume_state"):
            base = path.get("experiments_folder_path") or ""
            self.experiments_root = os.path.join(base, "experiments", "%s_%s" % (self.params["name"], get_current_datetime()))
        else:
            self.experiments_root = "/".join(path["resume_state"].split("/")[:-2])
        for key, val in list(path.items()):
            if not key.startswith("resume") and not key.startswith("experiments"):
                path[key] = os.path.join(self.experiments_root, val)
                mkdirs(path[key])
        path["experiments_root"] = self.experiments_root

    def __getattr__(self, item):
        return None


def dict2str(opt, indent_l=1):
    msg = ""
    for k, v in opt.items():
        if isinstance(v, dict):
            msg += " " * (indent_l * 2) + k + ":[\n" + dict2str(v, indent_l + 1) + " " * (indent_l * 2) + "]\n"
        else:
            msg += " " * (indent_l * 2) + k + ": " + str(v) + "\n"
    return msg
