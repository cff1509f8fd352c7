from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)   # the other GLIGEN modalities resolve to the reference's copies
