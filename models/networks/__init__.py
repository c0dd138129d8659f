"""Network factory (reference models/networks/__init__.py:8-43): ``define_G(opt)`` resolves
``opt.netG + 'generator'`` case-insensitively among the classes of models.networks.generator."""
import torch

import util.util as util
from models.networks.base_network import BaseNetwork
from models.networks.generator import *  # noqa: F401,F403  (DeepFillC2Generator, MDGenerator)


def find_network_using_name(target_network_name, filename):
    cls = util.find_class_in_module(target_network_name + filename, "models.networks." + filename)
    assert issubclass(cls, BaseNetwork), "Class %s should be a subclass of BaseNetwork" % cls
    return cls


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    return find_network_using_name(opt.netG, "generator").modify_commandline_options(parser, is_train)


def create_network(cls, opt):
    net = cls(opt)
    net.print_network()
    if len(opt.gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.cuda()
    if opt.init_type is not None:
        net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_G(opt):
    return create_network(find_network_using_name(opt.netG, "generator"), opt)
