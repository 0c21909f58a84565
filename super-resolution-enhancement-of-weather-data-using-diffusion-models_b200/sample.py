"""sample.py -- same command line as the reference's sample.py:26-39 for the accelerated path:

    python sample.py -c CONFIG.json -p MODEL_PATH -o OUT_DIR -gpu 0 [-n N]
    torchrun --nproc-per-node 8 sample.py -c CONFIG.json -o OUT_DIR -gpu 0,1,2,3,4,5,6,7     # batch-sharded

Builds the model through the reference's registry path (Config -> create_model -> DDPM -> define_diffusion),
``prepare_to_eval``, ``feed_data``, ``generate_sr``; writes the super-resolved fields as ``sr.npy`` (the reference's
cartopy plots, training/visualization.py, are outside the accelerated path).  ``-p`` may be omitted to sample from
random-init weights (smoke / benchmarking)."""
import argparse
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import wsr  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", type=str, required=True, help="JSON file for configuration")
    ap.add_argument("-p", "--model_path", type=str, default=None, help="Path to trained model")
    ap.add_argument("-o", "--output_path", type=str, required=True, help="Path to save output")
    ap.add_argument("-gpu", "--gpu_ids", type=str, default=None, help="GPU ids to use")
    ap.add_argument("-n", "--number_of_samples", type=int, default=1)
    ap.add_argument("-t", "--image_types", nargs="+", default=["SR"], choices=["HR", "SR", "LR", "INTERPOLATED", "DELTA", "AE"])
    ap.add_argument("-m", "--color_map", type=str, default="coolwarm", choices=["coolwarm", "heat_muted"])
    ap.add_argument("-d", "--date", type=str, default=None)
    args = ap.parse_args(argv)

    Config = wsr.sub("configs.config").Config
    create_model = wsr.sub("models.base_model").create_model
    data = wsr.sub("data_synthetic")
    par = wsr.sub("parallel")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    opt = Config(args, experiment=False).params
    opt["phase"] = "val"
    if args.model_path:
        opt["path"]["resume_state"] = args.model_path
    elif opt["path"].get("resume_state") and not os.path.exists(str(opt["path"]["resume_state"]) + "_gen.pth"):
        opt["path"]["resume_state"] = None                # config points at the authors' machine: random-init weights
    if opt["model"].get("pretrained_model", {}).get("model_path") and not os.path.exists(opt["model"]["pretrained_model"]["model_path"]):
        opt["model"]["pretrained_model"]["model_path"] = None
    model = create_model(opt, None)
    model.prepare_to_eval()
    handler = None
    if data.is_store(str(opt["data"].get("dataroot", ""))):
        # reference sample.py:45-56,76-79: -d DATE restricts the validation range to that hour and fetches it by date
        if args.date:
            from datetime import datetime, timedelta
            fmt = "%Y-%m-%d-%H"
            opt["data"]["months_subset"] = [datetime.strptime(args.date, fmt).month]
            opt["data"]["transform_groups"] = [opt["data"]["months_subset"]]
            opt["data"]["val_min_date"] = args.date
            opt["data"]["val_max_date"] = (datetime.strptime(args.date, fmt) + timedelta(hours=1)).strftime(fmt)
        handler = data.store_handler(opt, val_only=True)
    if handler is not None and args.date:
        batch, months = handler.get_data_by_date(args.date)
    else:
        batch, months = next(iter(data.batches_from_opt(opt, "val")))
    n_total = batch["SR"].shape[0]
    if world > 1:
        batch = par.shard_batch(batch, rank, world)
    model.feed_data((batch, months))
    model.generate_sr()
    sr = model.SR
    if world > 1:
        sr = par.gather_batch(sr, n_total)
    if rank == 0:
        os.makedirs(args.output_path, exist_ok=True)
        np.save(os.path.join(args.output_path, "sr.npy"), sr.float().cpu().numpy())
        print("saved %s %s" % (os.path.join(args.output_path, "sr.npy"), tuple(sr.shape)))
        if handler is not None:
            phys = handler.get_data_transformer().inverse_transform({"SR": sr}, months[:sr.shape[0]] if len(months) >= sr.shape[0] else months)
            np.save(os.path.join(args.output_path, "sr_physical.npy"), phys["SR"].float().cpu().numpy())
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
