"""Test stand-in for `omegaconf` (absent offline): the callers only use OmegaConf.create(dict) -> attribute access."""


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return _Cfg(v) if isinstance(v, dict) else v

    def __setattr__(self, k, v):
        self[k] = v


class OmegaConf:
    @staticmethod
    def create(d=None):
        return _Cfg(d or {})

    @staticmethod
    def load(path):
        raise NotImplementedError("stub")
