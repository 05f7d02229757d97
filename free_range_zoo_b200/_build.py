"""In-tree build of libfrz.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')


def build(verbose: bool = False, clean: bool = False) -> str:
    """Run ``make`` in csrc/ and return the path of the shared library."""
    if clean:
        subprocess.run(['make', '-C', CSRC, 'clean'], check=True, capture_output=not verbose)
    result = subprocess.run(['make', '-C', CSRC, '-j4'], capture_output=True, text=True)
    if verbose or result.returncode != 0:
        print(result.stdout)
        print(result.stderr)
    if result.returncode != 0:
        raise RuntimeError('building libfrz.so failed')
    return os.path.join(CSRC, 'libfrz.so')


if __name__ == '__main__':
    print(build(verbose=True))
