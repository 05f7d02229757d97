"""Recipe for ``oracle/_ref``: a byte-for-byte copy of the reference's Python package, so that the reference's own CPU
step path can be TIMED on the GPU box (``bench.py --impl reference`` / ``cpu_baseline``), where ``/root/reference``
does not exist.  TEST / BASELINE INFRASTRUCTURE: nothing under ``free_range_zoo_b200`` imports it.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is present)

``pip install --no-index --no-build-isolation --no-deps --target ... /root/reference`` was tried first and fails here:
the project's build backend (poetry-core) is not in the offline wheelhouse (``ModuleNotFoundError: No module named
'poetry'``).  The package is pure Python, so the recipe copies ``free_range_zoo/`` (minus image assets and caches) into
``oracle/_ref/free_range_zoo``.  ``oracle/_ref/`` is listed in .gitignore -- reference sources never enter the history
-- but not in .gpurunignore, so the copy travels with the snapshot like the built ``libfrz.so``.  Its third-party
imports that carry no transition arithmetic (tensordict, gymnasium, pettingzoo, free_range_rust, supersuit, sqlalchemy,
pygame) are served by the stand-ins of ``oracle/ref_shim.py``.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = '/root/reference/free_range_zoo'
TARGET = os.path.join(HERE, '_ref', 'free_range_zoo')


def build_ref(verbose: bool = False):
    """Copy the reference package when its checkout is present; returns the path of the copy, or None when there is
    neither a checkout nor an earlier copy."""
    if os.path.isdir(SOURCE):
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        shutil.copytree(SOURCE, TARGET, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.png', '*.jpg', '*.gif',
                                                                      '*.ttf', 'assets'))
        if verbose:
            count = sum(len(files) for _, _, files in os.walk(TARGET))
            print(f'copied {count} files of the reference package to {TARGET}')
    return TARGET if os.path.isfile(os.path.join(TARGET, '__init__.py')) else None


if __name__ == '__main__':
    path = build_ref(verbose=True)
    sys.exit(0 if path else 1)
