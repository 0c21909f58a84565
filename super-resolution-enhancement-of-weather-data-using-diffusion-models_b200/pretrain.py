"""pretrain.py -- prior pre-training with the reference's command line (pretrain.py:186-279: ``-c CONFIG -p {train,val} -gpu IDS``;
SURVEY.md 8f N4).

``model.name = "SimpleSR"``: the SimpleCNN prior trained with ``image_compare_loss`` (0.2 * FFT-MSE + 0.1 * 4-level Haar-MSE, ONE
kernel for value and gradient), hand-written backward through its three convolutions, Adam -- per epoch: a pass over the training
loader, then validation with the error metrics in physical units (reference ``train`` / ``evaluate``, :19-104; PSNR / SSIM are
torcheval / skimage wrappers outside the accelerated path).  ``model.name = "RRDBNet"``: the RRDB encoder trained with ``F.l1_loss``
on its SR image; its backward pass (``_RRDBPlan.backward``) runs the dense blocks on one gradient buffer that mirrors the forward
concat buffer.
Data comes from the on-disk store through ``DataHandler`` (same call as the reference), or synthetic batches when ``dataroot`` is
not a store."""
import argparse
import logging
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import wsr  # noqa: E402


def get_model(opt):
    """reference pretrain.py:143-166."""
    m = opt["model"]
    if m["name"] == "SimpleSR":
        model = wsr.sub("models.simple_cnn.Simple_CNN").SimpleCNN(scale_factor=4, channels=m["in_channel"])
        return model, wsr.sub("models.simple_cnn.loss").image_compare_loss
    if m["name"] == "RRDBNet":
        model = wsr.sub("models.rrdb_encoder.RRDBNet").RRDBNet(in_nc=m["in_channel"], out_nc=m["out_channel"], nf=m["hidden_size"],
                                                               nb=m["num_block"], gc=m["hidden_size"] // 2)
        return model, torch.nn.functional.l1_loss
    raise ValueError(f"Unknown model name: {m['name']}")


def train_epoch(model, loader, criterion, optimizer, device):
    """reference :19-53; the loss values stay on the device until the end of the epoch (no per-iteration host sync)."""
    model.train()
    total = torch.zeros((), device=device)
    n = 0
    for batch, _ in loader:
        outputs = model(batch["LR"].to(device))
        loss = criterion(outputs, batch["HR"].to(device))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        total += loss.detach()
        n += 1
    return float(total) / max(n, 1), n


@torch.no_grad()
def evaluate(model, loader, device, data_transformer=None, sigma=1.0):
    """reference :56-104 with the device-side accumulators: MSE / RMSE / MAE / MR of the prior's output against HR, in physical
    units (per-sample fitted std folded into the metric pass) when a fitted transformer is available."""
    metrics = wsr.sub("training.metrics")
    model.eval()
    vm = metrics.ValidationMetrics(metrics.create_metric_dict(device))
    for batch, months in loader:
        out = model(batch["LR"].to(device))
        hr = batch["HR"].to(device)
        if data_transformer is not None:
            _, std = data_transformer.batch_statistics("hr", months)
            vm.update(out, hr, scale=std.reshape(-1))
        else:
            vm.update(out, hr, scale=torch.full((out.shape[0] * out.shape[1],), float(sigma)))
    return vm.compute_metrics()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    ap.add_argument("-p", "--phase", type=str, choices=["train", "val"], default="train")
    ap.add_argument("-gpu", "--gpu_ids", type=str, default=None)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    log = logging.getLogger("base")
    import random
    import numpy as np
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    opt = wsr.sub("configs.config").Config(args, experiment=True).params
    device = torch.device("cuda", 0)
    data = wsr.sub("data_synthetic")
    handler = data.store_handler(opt)
    transformer = handler.get_data_transformer() if handler is not None else None

    def loader(phase, n=None):
        if handler is not None:
            return handler.train_loader if phase == "train" else handler.val_loader
        d = opt["data"]
        m = opt["model"]
        bs = d["val_batch_size"] if phase == "val" else d["batch_size"]
        return list(data.synthetic_batches(n or 4, bs, m["in_channel"], d.get("height", 128), 2 * d.get("height", 128)))

    model, criterion = get_model(opt)
    model = model.to(device)
    if opt["path"].get("resume_state") and os.path.exists(str(opt["path"]["resume_state"])):
        log.info("Loading pretrained model [%s]", opt["path"]["resume_state"])
        model.load_state_dict(torch.load(opt["path"]["resume_state"], map_location=device))
    if opt["phase"] == "val":
        res = evaluate(model, loader("val"), device, transformer)
        log.info("Val " + ", ".join("%s: %.4f" % (k, float(v)) for k, v in res.items()))
        return res
    oc = opt["train"]["optimizer"]
    FusedAdam = wsr.sub("autograd_glue").FusedAdam
    if oc.get("amsgrad"):
        raise NotImplementedError("amsgrad is not supported by the fused Adam step")
    optimizer = FusedAdam(model.parameters(), lr=oc["lr"])
    history = []
    for epoch in range(opt["train"]["epoch"]):
        train_loss, n_iter = train_epoch(model, loader("train"), criterion, optimizer, device)
        res = evaluate(model, loader("val"), device, transformer)
        history.append((train_loss, {k: float(v) for k, v in res.items()}))
        log.info("Epoch [%d/%d], Iter %d, Train Loss: %.4f, Val RMSE: %.4f, MSE: %.4f, MAE: %.4f", epoch + 1, opt["train"]["epoch"],
                 n_iter, train_loss, float(res["RMSE"]), float(res["MSE"]), float(res["MAE"]))
        ckpt = opt["path"].get("checkpoint")
        if ckpt:
            os.makedirs(ckpt, exist_ok=True)
            name = opt.get("diffusion", {}).get("name", opt["model"]["name"]) if isinstance(opt.get("diffusion"), dict) else opt["model"]["name"]
            torch.save(model.state_dict(), os.path.join(ckpt, f"pretrain_{name}_E{epoch}_gen.pth"))
    return history


if __name__ == "__main__":
    main()
