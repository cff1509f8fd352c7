from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)   # modules not defined here resolve to the reference's copies
