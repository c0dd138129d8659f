"""Model factory with the reference's surface: ``models.create_model(opt)`` resolves
``--model <name>`` to the ``<Name>Model`` class in ``models/<name>_model.py``
(reference models/__init__.py:5-39). Only ``editline2`` exists on this path."""
import importlib

import torch


def find_model_using_name(model_name):
    module = importlib.import_module("models.%s_model" % model_name)
    wanted = (model_name.replace("_", "") + "model").lower()
    for attr, obj in vars(module).items():
        if attr.lower() == wanted and isinstance(obj, type) and issubclass(obj, torch.nn.Module):
            return obj
    raise SystemExit("models/%s_model.py defines no torch.nn.Module subclass named like %r" % (model_name, wanted))


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(instance).__name__)
    return instance
