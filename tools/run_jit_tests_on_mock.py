"""Run EVERY GPU test of tests/test_jit_gpu.py (all parametrisations) against the mock engine
(tests/_mock_engine.py) on the CPU.  The CPU suite runs a subset (tests/test_host_logic_mock.py);
this is the exhaustive version:  python tools/run_jit_tests_on_mock.py"""
import os, sys, inspect, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pytest
import _mock_engine as me
from oracle import oracle
oracle.build()
import test_jit_gpu as tj

class MP:
    def setenv(self, k, v):
        import os; os.environ[k]=v
fails=0
for name, fn in inspect.getmembers(tj, inspect.isfunction):
    if not name.startswith('test_'): continue
    marks=[m for m in getattr(fn,'pytestmark',[]) if m.name=='parametrize']
    combos=[{}]
    for m in marks:
        names=[n.strip() for n in m.args[0].split(',')]
        new=[]
        for c in combos:
            for vals in m.args[1]:
                vals = vals if isinstance(vals,(tuple,list)) and len(names)>1 else (vals,)
                d=dict(c); d.update(dict(zip(names,vals))); new.append(d)
        combos=new
    for c in combos:
        with me.install(oracle) as eng:
            kw=dict(c)
            for p in inspect.signature(fn).parameters:
                if p=='engine': kw[p]=eng
                elif p=='oracle': kw[p]=oracle
                elif p=='monkeypatch': kw[p]=MP()
                elif p=='matrix_kernel': kw[p]=-1
            try:
                fn(**kw); print('PASS', name, c, flush=True)
            except Exception as e:
                fails+=1; print('FAIL', name, c, repr(e)[:200], flush=True); traceback.print_exc()
os.environ.pop('FDB_AFFINE',None)
print("failures", fails)
sys.exit(1 if fails else 0)
