"""TEST INFRASTRUCTURE — empty shim of `omegaconf` (absent from this image): only the names
/root/reference/tactile_ssl/utils/logging.py:8 imports; nothing on the VTT arithmetic path uses them."""


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True):
        return dict(cfg)

    @staticmethod
    def to_yaml(cfg, resolve=True):
        return str(dict(cfg))
