"""BaseModel / create_model -- drop-in for the reference's models/base_model.py:7-141."""
import typing
from abc import abstractmethod

import torch
import torch.nn as nn


class BaseModel:
    def __init__(self, opt):
        self.opt = opt
        self.device = torch.device("cuda" if torch.cuda.is_available() and opt['gpu_ids'] else "cpu")
        self.begin_step, self.begin_epoch = 0, 0

    @abstractmethod
    def feed_data(self, data) -> None:
        pass

    @abstractmethod
    def optimize_parameters(self) -> None:
        pass

    @abstractmethod
    def get_images(self) -> dict:
        pass

    @abstractmethod
    def print_network(self) -> None:
        pass

    def set_device(self, x):
        if isinstance(x, dict):
            return {k: (v.to(self.device) if v.numel() else v) for k, v in x.items()}
        if isinstance(x, list):
            return [v.to(self.device) if v else v for v in x]
        return x.to(self.device)

    @staticmethod
    def get_network_description(network: nn.Module) -> typing.Tuple[str, int]:
        if isinstance(network, nn.DataParallel):
            network = network.module
        return str(network), sum(p.numel() for p in network.parameters())

    def get_loaded_epoch(self):
        return self.begin_epoch

    def get_loaded_iter(self):
        return self.begin_step

    @abstractmethod
    def prepare_to_train(self) -> None:
        pass

    @abstractmethod
    def prepare_to_eval(self) -> None:
        pass

    @abstractmethod
    def generate_sr(self, continuous: bool = False) -> None:
        pass

    @abstractmethod
    def save_network(self):
        pass

    @abstractmethod
    def get_current_log(self) -> dict:
        pass


def create_model(opt, optimizer=None):
    if opt["model"]["model_name"] == "diffusion":
        from . import diffusion_models as model
        return model.create_model(opt)
    raise NotImplementedError("Model {} not implemented.".format(opt["model"]["model_name"]))
