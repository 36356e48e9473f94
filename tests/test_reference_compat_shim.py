"""baseline/compat/sitecustomize.py: the interpreter-start shim for the UNMODIFIED reference's AD-PSGD
gossip process (a forkserver child that reads the pre-c10d attribute `torch.distributed._backend`,
gossip/gossiper.py:50-51).  Checked here without a GPU: a forkserver child started with the shim on
PYTHONPATH sees the legacy UNDEFINED value (-1, which passes the reference's two asserts), and the shim
chains to another sitecustomize further down the path instead of shadowing it."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, 'baseline', 'compat')


def test_forkserver_child_sees_legacy_backend_attribute(tmp_path):
    script = tmp_path / 'fs.py'
    script.write_text(textwrap.dedent('''
        import os
        import multiprocessing as mp

        def child(q):
            import torch.distributed as dist
            q.put(getattr(dist, '_backend', 'MISSING'))

        if __name__ == '__main__':
            os.environ['PYTHONPATH'] = %r + os.pathsep + os.environ.get('PYTHONPATH', '')
            mp.set_start_method('forkserver')
            q = mp.Queue()
            p = mp.Process(target=child, args=(q,))
            p.start()
            print('BACKEND', q.get(timeout=240))
            p.join()
    ''' % COMPAT))
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    out = subprocess.run([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:]
    assert 'BACKEND -1' in out.stdout, out.stdout[-2000:]


def test_shim_chains_to_the_next_sitecustomize(tmp_path):
    other = tmp_path / 'other'
    other.mkdir()
    (other / 'sitecustomize.py').write_text("import os\nos.environ['OTHER_HOOK_RAN'] = '1'\n")
    env = dict(os.environ)
    env['PYTHONPATH'] = COMPAT + os.pathsep + str(other)
    out = subprocess.run([sys.executable, '-c',
                          "import os, torch.distributed as d; print(d._backend, os.environ.get('OTHER_HOOK_RAN'))"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.strip().endswith('-1 1'), out.stdout[-2000:]
