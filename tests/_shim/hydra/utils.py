def instantiate(cfg, *a, **k):
    raise NotImplementedError("hydra is not available; construct objects with plain kwargs")
