# -*- coding: utf-8 -*-
"""Configuration and logging helpers mirroring mtscomp's (reference mtscomp.py:46-108, 176-209)."""

import json
import logging
import multiprocessing as mp
import os.path as op
import sys
from pathlib import Path

FORMAT_VERSION = '1.0'
CHECK_ATOL = 1e-16

# Same keys and defaults as the reference's DEFAULT_CONFIG (mtscomp.py:46-57), kept read-only as a tuple of pairs.
DEFAULT_CONFIG = (
    ('algorithm', 'zlib'),
    ('cache_size', 10),
    ('check_after_compress', True),
    ('check_after_decompress', True),
    ('chunk_duration', 1.),
    ('chunk_order', 'F'),
    ('comp_level', -1),
    ('do_spatial_diff', False),
    ('do_time_diff', True),
    ('n_threads', mp.cpu_count()),
)

logger = logging.getLogger('mtscomp')
logger.setLevel(logging.INFO)
logger.addHandler(logging.NullHandler())


class Bunch(dict):
    """dict with attribute access (reference mtscomp.py:99-104)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class _ColorFormatter(logging.Formatter):
    _colors = {'D': '90', 'I': '0', 'W': '33', 'E': '31'}

    def format(self, record):
        record.levelname = record.levelname[0]
        stem = op.splitext(op.basename(record.pathname))[0]
        record.caller = ('%s:%d' % (stem, record.lineno)).ljust(20)
        text = super().format(record)
        return '\33[%sm%s\33[0m' % (self._colors.get(record.levelname, '7'), text)


def add_default_handler(level='INFO', logger=logger):
    """Attach a coloured stream handler (reference mtscomp.py:89-96)."""
    h = logging.StreamHandler()
    h.setLevel(level)
    h.setFormatter(_ColorFormatter(
        fmt='%(asctime)s.%(msecs)03d [%(levelname)s] %(caller)s %(message)s', datefmt='%H:%M:%S'))
    logger.addHandler(h)


def config_path():
    return (Path('~') / '.mtscomp').expanduser()


def _current_config_path():
    # tests (and users) may rebind mtscomp_b200.CONFIG_PATH, as the reference's tests do with mtscomp.CONFIG_PATH
    pkg = sys.modules.get('mtscomp_b200')
    return Path(getattr(pkg, 'CONFIG_PATH', None) or config_path())


def read_config(**kwargs):
    """defaults <- ~/.mtscomp JSON <- non-None kwargs (reference mtscomp.py:186-200)."""
    params = dict(DEFAULT_CONFIG)
    path = _current_config_path()
    user = {}
    if path.exists():
        with path.open('r') as f:
            user = json.load(f)
    for layer in (user, kwargs):
        for k, v in layer.items():
            if v is not None:
                params[k] = v
    return Bunch(params)


def write_config(**kwargs):
    """Persist the merged configuration (reference mtscomp.py:203-209)."""
    cfg = read_config(**kwargs)
    path = _current_config_path()
    path.parent.mkdir(exist_ok=True, parents=True)
    with path.open('w') as f:
        json.dump(cfg, f, indent=2, sort_keys=True)
    return cfg


def clip(x, lo, hi):
    return max(lo, min(hi, x))
