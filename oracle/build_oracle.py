"""Builds the C restatement (oracle/mmi_oracle.c) into oracle/_build/libmmi_oracle.so and the host-compiled copy of
the input-stage pixel function and of the B2A output function (oracle/preprocess_host.cpp, oracle/attributes_host.cpp, oracle/metrics_host.cpp)
into oracle/_build/.

TEST INFRASTRUCTURE ONLY.  Building the checker is not using it: __graft_entry__.build()
calls this so the .so travels to the GPU box with the snapshot.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, 'mmi_oracle.c')
    out = os.path.join(out_dir, 'libmmi_oracle.so')
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', out,
                               src, '-lm'])
    build_preprocess_host(force)
    build_attributes_host(force)
    build_metrics_host(force)
    return out


def build_preprocess_host(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, 'preprocess_host.cpp')
    hdr = os.path.join(os.path.dirname(_HERE), 'shapy_b200', 'csrc', 'preprocess.cuh')
    out = os.path.join(out_dir, 'libpreprocess_host.so')
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', out, src])
    return out


def build_attributes_host(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, 'attributes_host.cpp')
    hdr = os.path.join(os.path.dirname(_HERE), 'shapy_b200', 'csrc', 'attributes.cuh')
    out = os.path.join(out_dir, 'libattributes_host.so')
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', out, src])
    return out


def build_metrics_host(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, 'metrics_host.cpp')
    hdr = os.path.join(os.path.dirname(_HERE), 'shapy_b200', 'csrc', 'metrics.cuh')
    out = os.path.join(out_dir, 'libmetrics_host.so')
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', out, src])
    return out


if __name__ == '__main__':
    print(build(force=True))
