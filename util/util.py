"""Checkpoint I/O and class lookup with the reference's behaviour (reference util/util.py:175-225)."""
import importlib
import os

import torch


def find_class_in_module(target_cls_name, module):
    wanted = target_cls_name.replace("_", "").lower()
    lib = importlib.import_module(module)
    for name, obj in vars(lib).items():
        if name.lower() == wanted:
            return obj
    raise SystemExit("In %s, there should be a class whose name matches %s in lowercase without underscore(_)"
                     % (module, wanted))


def _strip_module_prefix(weights):
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in weights.items()}


def load_network(net, label, epoch, opt):
    """<checkpoints_dir>/<name>/<epoch>_net_<label>.pth, 'module.' prefix stripped, strict load."""
    path = os.path.join(opt.checkpoints_dir, opt.name, "%s_net_%s.pth" % (epoch, label))
    net.load_state_dict(_strip_module_prefix(torch.load(path, map_location="cpu")))
    return net


def load_network_path(net, save_path):
    net.load_state_dict(_strip_module_prefix(torch.load(save_path, map_location="cpu")), strict=False)
    return net


def save_network(net, label, epoch, opt):
    path = os.path.join(opt.checkpoints_dir, opt.name, "%s_net_%s.pth" % (epoch, label))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({k: v.cpu() for k, v in net.state_dict().items()}, path)
