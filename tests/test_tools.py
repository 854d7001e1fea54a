"""Small host-side tools (no GPU)."""
import os
import subprocess
import sys

from conftest import ROOT


def test_ncu_summary_condenses_a_raw_export(tmp_path):
    raw = tmp_path / 'raw.csv'
    raw.write_text(
        '"ID","Kernel Name","Grid Size","SM_A.X.gpu__time_duration.sum","launch__grid_size","dram__bytes_read.sum"\n'
        '"","","","us","","Mbyte"\n'
        '"0","void dccf::k_a<2>(dccf::P)","1","10.0","148","2.0"\n'
        '"1","dccf::k_b(const float *, int)","1","4.0","16","0.5"\n'
        '"2","void dccf::k_a<2>(dccf::P)","1","14.0","148","4.0"\n')
    tool = os.path.join(ROOT, 'tools', 'ncu_summary.py')
    r = subprocess.run([sys.executable, tool, str(raw), '-c', 'hello'], capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0] == '# hello'
    assert lines[1] == 'Kernel Name,gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum'
    assert lines[2] == ',us,,Mbyte'
    assert lines[3:] == ['k_a<2>,10.0,148,2.0', 'k_b,4.0,16,0.5', 'k_a<2>,14.0,148,4.0']
    r = subprocess.run([sys.executable, tool, str(raw), '--mean', '-k', 'k_a'], capture_output=True, text=True, check=True)
    assert r.stdout.strip().splitlines()[-1] == 'k_a<2>,12,148,3,2'
