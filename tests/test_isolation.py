"""The oracle and the reference checkout are test infrastructure: the product package never imports `oracle`, and nothing that runs on
the GPU box (the package, bench.py, __graft_entry__, the `-m gpu` tests) reads /root/reference."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _code(path):
    """Source with comments and docstrings left in: a mention in prose is harmless, so only import / path-use lines are matched."""
    return open(path).read()


def test_product_package_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "vit_deep_radiomics_b200", "*.py")):
        assert not re.search(r"^\s*(from\s+oracle\b|import\s+oracle\b|from\s+\.\.?oracle\b)", _code(path), flags=re.M), path
        assert "ref_shim" not in _code(path), path
    for path in glob.glob(os.path.join(ROOT, "vit_deep_radiomics_b200", "csrc", "*")):
        if path.endswith((".cu", ".cuh", ".h", "Makefile")):
            assert not re.search(r"#\s*include[^\n]*oracle|-I[^\n]*oracle|oracle/[A-Za-z_]+\.(c|o|so)\b", _code(path)), path   # comments may cite it


def test_bench_uses_the_oracle_only_in_its_cpu_legs_and_parity_check():
    src = _code(os.path.join(ROOT, "bench.py"))
    funcs = re.split(r"^def ", src, flags=re.M)
    for f in funcs[1:]:
        name = f.split("(", 1)[0]
        if re.search(r"^\s+from oracle import", f, flags=re.M):
            # every function that touches the oracle is a CPU leg / parity check by name or says so in its docstring
            head = f[:1500].lower()
            assert any(k in name.lower() for k in ("cpu", "reference", "parity", "baseline")) or "cpu" in head or "oracle" in head, name


def test_nothing_on_the_gpu_box_reads_the_reference_checkout():
    paths = (glob.glob(os.path.join(ROOT, "vit_deep_radiomics_b200", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py"))
             + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    for path in paths:
        for line in _code(path).splitlines():
            code = line.split("#", 1)[0]
            if "/root/reference" in code:
                assert re.search(r'"""|\'\'\'|^\s*[A-Za-z(`]', code) and "open(" not in code and "sys.path" not in code \
                    and "import" not in code, (path, line)
