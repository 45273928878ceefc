"""TEST INFRASTRUCTURE ONLY: imports the UNMODIFIED reference from /root/reference on CPU.

Only usable in the authoring container (the GPU box has no /root/reference).  It is used by
oracle/gen_golden.py to produce tests/golden/*.npz and by tests that pin the oracle restatement
against the live reference when it is present.  Nothing in como_b200/ imports this.

What it takes: two import shims (oracle/shims: lietorch, pytorch_lightning)
and a CPU build of the reference's own pybind module `como_backends` compiled from the sources
where they lie (cov.cpp, cov_cpu.cpp, depth_cov_backends.cpp) into oracle/_ref/.
"""
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "como"))


def build_ref_backends(verbose=False):
    """Compile the reference's CPU `como_backends` into oracle/_ref (outputs only there)."""
    from torch.utils.cpp_extension import load

    os.makedirs(REF_BUILD, exist_ok=True)
    R = os.path.join(REF, "como", "backend")
    return load(
        name="como_backends",
        sources=[f"{R}/src/cov.cpp", f"{R}/src/cov_cpu.cpp", f"{R}/src/depth_cov_backends.cpp"],
        extra_include_paths=[f"{R}/include", "/usr/local/cuda/include"],
        extra_cflags=["-O3"],
        build_directory=REF_BUILD,
        verbose=verbose,
    )


_loaded = False


def load_reference():
    """Put the reference + shims on sys.path; returns the `como` package."""
    global _loaded
    if not available():
        raise RuntimeError("reference not present (this only works in the authoring container)")
    if not _loaded:
        build_ref_backends()
        sys.path[:0] = [os.path.join(HERE, "shims"), REF_BUILD, REF]
        _loaded = True
    import como  # noqa

    return como
