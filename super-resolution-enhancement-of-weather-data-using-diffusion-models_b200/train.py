"""train.py -- same command line as the reference's train.py:208-212 (``-c CONFIG -p {train,val} -gpu IDS``).

``-p val`` runs the validation loop (generate_sr per batch, MSE / RMSE / MAE / MR on device accumulators) on the CUDA path.
``-p train`` builds the optimiser and runs ``optimize_parameters`` (forward, hand-written backward, fused Adam).  Under
``torchrun`` (one process per GPU) the batch is sharded across ranks and the gradients are all-reduced in buckets
overlapped with the backward pass (parallel.FlatGradReducer); weights start identical on every rank (same seed)."""
import argparse
import logging
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import wsr  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    ap.add_argument("-p", "--phase", type=str, choices=["train", "val"], default="train")
    ap.add_argument("-gpu", "--gpu_ids", type=str, default=None)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    log = logging.getLogger("base")

    Config = wsr.sub("configs.config").Config
    create_model = wsr.sub("models.base_model").create_model
    data = wsr.sub("data_synthetic")
    import numpy as np
    import random
    random.seed(0); np.random.seed(0); torch.manual_seed(0)                 # training/utils.py:39-50
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        args.gpu_ids = os.environ.get("LOCAL_RANK", "0")
    opt = Config(args, experiment=True).params
    for key in ("resume_state",):
        if opt["path"].get(key) and not os.path.exists(str(opt["path"][key]) + "_gen.pth"):
            opt["path"][key] = None
    pm = opt["model"].get("pretrained_model", {})
    if pm.get("model_path") and not os.path.exists(pm["model_path"]):
        pm["model_path"] = None
    model = create_model(opt, None)
    if args.phase == "val":
        # reference train.py:132-196 (validate): generate_sr per batch, inverse transform, metric accumulators -- here the
        # accumulators live on the device (training/metrics.py) and the inverse StandardScaling is folded into the same pass
        # as a per-(sample, variable) scale, so nothing is copied to the host until the final compute
        model.prepare_to_eval()
        metrics = wsr.sub("training.metrics")
        vm = metrics.ValidationMetrics(metrics.create_metric_dict(model.device))
        vm_k = metrics.ValidationMetrics(metrics.create_metric_dict(model.device))
        sigma_k = float(opt["data"].get("sigma_kelvin", 21.26))          # WeatherBench t2m std (synthetic data has no fitted transform)
        handler = data.store_handler(opt)                                # the reference's on-disk store, when dataroot is one
        for batch, months in data.batches_from_opt(opt, "val"):
            model.feed_data((batch, months))
            model.generate_sr(False)
            sr, hr = model.SR, model.data["HR"]
            vm.update(sr, hr)
            if handler is not None:
                # physical units = fitted std of each sample's month and variable (train.py:96-99 inverse_transform, folded in)
                _, std = handler.get_data_transformer().batch_statistics("hr", months)
                vm_k.update(sr, hr, scale=std.reshape(-1))
            else:
                vm_k.update(sr, hr, scale=torch.full((sr.shape[0] * sr.shape[1],), sigma_k))
        vm.compute_metrics(); vm_k.compute_metrics()
        log.info("validation (standardised units)%s", vm.metrics2str())
        if handler is not None:
            log.info("validation (physical units, fitted statistics)%s", vm_k.metrics2str())
        else:
            log.info("validation (x sigma = %.2f K)%s", sigma_k, vm_k.metrics2str())
        return vm.metrics2dict(), vm_k.metrics2dict()
    it = 0
    par = wsr.sub("parallel")
    from_store = data.is_store(str(opt["data"].get("dataroot", "")))
    shard = (rank, world) if (world > 1 and from_store) else None        # the store loader reads only this rank's slice
    for batch, months in data.batches_from_opt(opt, "train", shard=shard):
        it += 1
        if world > 1 and shard is None:
            batch = par.shard_batch(batch, rank, world)
        model.feed_data((batch, months))
        model.optimize_parameters()
        if it % opt["train"]["print_freq"] == 0:
            log.info("iter %d  l_pix %.6f", it, model.get_current_log()["l_pix"])
        if it % opt["train"]["save_checkpoint_freq"] == 0 and rank == 0:
            model.save_network(0, it)


if __name__ == "__main__":
    main()
