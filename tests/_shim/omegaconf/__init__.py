import contextlib


class DictConfig(dict):
    pass


class OmegaConf:
    to_yaml = staticmethod(str)


def open_dict(cfg):
    return contextlib.nullcontext()
