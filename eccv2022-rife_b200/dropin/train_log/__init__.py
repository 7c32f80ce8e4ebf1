# Namespace-style package so that the reference's other `train_log.*` modules (e.g. refine.py) stay importable
# when the reference directory is also on sys.path.
import pkgutil
__path__ = pkgutil.extend_path(__path__, __name__)
