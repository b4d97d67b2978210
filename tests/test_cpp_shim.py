"""The C++20 host shim (include/Einsums/...): its compile-time analyser is checked on the CPU against the golden
AlgorithmChoice list extracted from the reference's tests; its device path runs the reference's GPU test cases on the B200."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
OUT = os.path.join(ROOT, "build_cpp")


def _make():
    out = subprocess.run(["make", "-C", CPP, "-j", "4"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


def test_analyser_static_asserts_compile_and_run(eb):
    _make()
    out = subprocess.run([os.path.join(OUT, "test_analysis")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all passed at compile time" in out.stdout


def test_analyser_against_golden_json(eb):
    """Generate one static_assert per golden entry (tests/golden/algorithm_choice.json) and compile it: a wrong decision
    of the C++ analyser is a compile error."""
    entries = json.load(open(os.path.join(ROOT, "tests", "golden", "algorithm_choice.json")))
    lines = ['#include <Einsums/B200/Analysis.hpp>', '#include <Einsums/B200/Index.hpp>', 'using namespace einsums;',
             'using namespace einsums::index;', 'using namespace einsums::tensor_algebra::detail;',
             'template <typename... Cs, typename... As, typename... Bs> constexpr AlgorithmChoice choose(std::tuple<Cs...> const &, std::tuple<As...> const &, std::tuple<Bs...> const &) {',
             '  return choice_of(static_route(make_pack(ids_of<std::tuple<Cs...>>()), make_pack(ids_of<std::tuple<As...>>()), make_pack(ids_of<std::tuple<Bs...>>()), false, false, true)); }']
    n = 0
    for e in entries:
        if e["source"].endswith("Einsum7.cpp:209"):
            continue  # run-time (strided view) decision, covered by test_device.cpp:gpu_views_runtime_decision
        pk = lambda p: "Indices{" + ", ".join(p) + "}"
        lines.append(f'static_assert(choose({pk(e["C"])}, {pk(e["A"])}, {pk(e["B"])}) == {e["expected"]}, "{e["source"]}");')
        n += 1
    lines.append("int main() { return 0; }")
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(OUT, "gen_golden.cpp")
    open(src, "w").write("\n".join(lines) + "\n")
    out = subprocess.run(["/usr/bin/g++", "-std=c++20", "-fsyntax-only", f"-I{ROOT}/include", src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
    assert n >= 40


@pytest.mark.gpu
def test_device_shim_cases(gpu):
    _make()
    out = subprocess.run([os.path.join(OUT, "test_device")], capture_output=True, text=True, timeout=600)
    print(out.stdout[-6000:])
    assert out.returncode == 0, out.stdout[-6000:] + out.stderr[-2000:]
    assert "0 failed checks" in out.stdout
