"""Drop-in portion of the reference's ``utils.brain_tumor_utils`` package (config_parser, logger, io).

The reference's ``__init__`` re-exports helpers from its sub-modules (reference: src/utils/brain_tumor_utils/__init__.py:1-4);
here they resolve lazily, so importing the package does not pull in ``datautils`` (PIL / torchvision) on the GPU hot
path.  Modules this repo does not replace (``datautils``) are found in the reference's own directory further down
``sys.path`` through ``pkgutil.extend_path``."""
import importlib
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

_EXPORTS = {
    "get_config": "config_parser",
    "init_logger": "logger", "log_config": "logger", "log_metrics": "logger",
    "ensure_dirs": "io", "save_checkpoint": "io", "save_json": "io", "save_table": "io", "save_figure": "io",
    "model_checkpoint_path": "io", "run_artifact_dir": "io",
    "save_sharded_checkpoint": "io", "load_sharded_checkpoint": "io",
    "BrainTumorDataset": "datautils", "build_dataloaders": "datautils",
}


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
