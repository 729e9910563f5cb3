"""Drop-in portion of the reference's ``inference`` package.

``pkgutil.extend_path`` lets the package span several ``sys.path`` entries: modules this repo replaces are found here
(this directory is first on the path), every other ``inference.*`` module resolves to the reference's own file further
down ``sys.path`` -- so the reference's scripts (``src/training/train.py`` and friends) run unmodified with the
B200 model underneath."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
