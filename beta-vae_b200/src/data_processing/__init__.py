# coexists with the reference's src/data_processing (its augmentations.py / preprocessing stay importable)
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
