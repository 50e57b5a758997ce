"""Race check without a GPU (tools/emu_order_check.py): the emulated kernels executed with blocks, threads and coroutines in REVERSE and in a pseudo-random order
must give the same bits as in forward order -- a kernel that reads what another thread or block of the same launch writes would not."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernels_do_not_depend_on_execution_order():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_order_check.py")], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("bit-identical") >= 7 and "DIFFERS" not in out.stdout, out.stdout
