"""CPU tests of the launcher for scripts with a pasted copy of the layer (graphlearninglayer_b200/run.py)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    """stand-in for train_and_adversarial.py: pasted class, module-level driver, optional imports"""
    import sys
    import graphlearning as gl          # not installed here
    import umap                          # not installed here
    import torch

    class LaplaceLearningSparseHard(torch.autograd.Function):   # the pasted copy (train_and_adversarial.py:26)
        PASTED = True
        @staticmethod
        def forward(ctx, X, label_matrix, tau=0, epsilon='auto'):
            raise RuntimeError("the pasted CPU class must not run")

    def knn_sym_dist(data, k=25, epsilon='auto'):               # the pasted helper (train_and_adversarial.py:202)
        raise RuntimeError("the pasted helper must not run")

    def train(lap):
        return lap.__module__

    lap = LaplaceLearningSparseHard
    print("ARGS", sys.argv[1:])
    print("CLASS", train(lap), getattr(lap, "PASTED", False))
    print("HELPER", knn_sym_dist.__module__)
''')


def test_rewrite_drops_pasted_definitions():
    sys.path.insert(0, ROOT)
    from graphlearninglayer_b200 import run

    code, dropped = run.rewrite(SCRIPT, "fake.py")
    assert any("ClassDef LaplaceLearningSparseHard" in d for d in dropped)
    assert any("FunctionDef knn_sym_dist" in d for d in dropped)
    assert len(dropped) == 2


def test_launcher_end_to_end(tmp_path):
    script = tmp_path / "train_and_adversarial.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run([sys.executable, "-m", "graphlearninglayer_b200.run", str(script), "gl", "natural", "mnist"],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=300)
    assert p.returncode == 0, p.stderr
    out = p.stdout
    assert "ARGS ['gl', 'natural', 'mnist']" in out                 # positional sys.argv driver (t_a_a.py:756-775)
    assert "CLASS graphlearninglayer_b200.GLL False" in out         # our Function, not the pasted one
    assert "HELPER graphlearninglayer_b200.GLL" in out
    assert "stand-in modules" in p.stderr and "graphlearning" in p.stderr
