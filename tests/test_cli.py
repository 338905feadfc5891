"""`bnz` CLI parity with the reference front end (bnz/src/main.rs): flags, messages that scripts
rely on, exit codes 0/1/2/3 (main.rs:11-14).  CPU tests cover argument handling; the GPU test
covers a real round trip including the default-delete rule (main.rs:292-309)."""
import bz2
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BNZ = os.path.join(ROOT, "banzai_b200", "bnz")


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(BNZ)


def run(*args, stdin=None):
    return subprocess.run([BNZ, *args], input=stdin, capture_output=True)


def test_no_arguments_prints_synopsis_and_exits_1():
    r = run()
    assert r.returncode == 1 and b"--help" in r.stderr


@pytest.mark.parametrize("flag", ["--help", "--info", "--version"])
def test_commands_exit_0_and_write_to_stderr(flag):
    r = run(flag)
    assert r.returncode == 0 and r.stdout == b"" and r.stderr


def test_argument_errors_exit_1():
    assert run("--bogus", "x").returncode == 1
    assert run("-x", "x").returncode == 1
    assert run("a", "b").returncode == 1                       # only one input
    assert run("--output", "-c", "x").returncode == 1          # --output needs a path
    assert run("--output", "a", "--output", "b", "x").returncode == 1
    assert run("-k").returncode == 1                           # no input
    assert b"Flag 'x' is not valid" in run("-kx", "y").stderr


def test_missing_input_file_exits_2(tmp_path):
    r = run(str(tmp_path / "does-not-exist"))
    assert r.returncode == 2 and b"[filesystem error]" in r.stderr


@pytest.mark.gpu
def test_round_trip_and_default_delete(tmp_path):
    data = (b"banzai on a B200, " * 40000)
    p = tmp_path / "in.txt"
    p.write_bytes(data)
    # explicit output -> input kept
    out = tmp_path / "explicit.bz2"
    r = run("-5", "--output", str(out), str(p))
    assert r.returncode == 0, r.stderr
    assert bz2.decompress(out.read_bytes()) == data and p.exists()
    assert out.read_bytes()[:4] == b"BZh5"
    # stdout + combined short flags
    r = run("-kc1", str(p))
    assert r.returncode == 0 and bz2.decompress(r.stdout) == data and r.stdout[:4] == b"BZh1"
    # stdin -> stdout
    r = run("-", stdin=data)
    assert r.returncode == 0 and bz2.decompress(r.stdout) == data
    # default: <input>.bz2 written, input removed
    r = run(str(p))
    assert r.returncode == 0 and not p.exists()
    assert bz2.decompress((tmp_path / "in.txt.bz2").read_bytes()) == data


@pytest.mark.gpu
def test_failing_output_keeps_the_input(tmp_path):
    """a sink that cannot take the bytes (/dev/full) is an output error (exit 3, main.rs:287-290) and
    the input is not removed, whatever the delete rule says"""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    data = os.urandom(300000)
    p = tmp_path / "in.bin"
    p.write_bytes(data)
    r = run("--remove", "--output", "/dev/full", str(p))
    assert r.returncode == 3, r.stderr
    assert p.exists() and p.read_bytes() == data
