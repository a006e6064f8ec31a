import torch


class Trainer:
    pass


def seed_everything(seed):
    torch.manual_seed(seed)
    return seed
