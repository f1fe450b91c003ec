import sys, json
sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np, torch, helpers as Hh
import g4splat_b200.diff_surfel_rasterization as op
from oracle.oracle import Oracle
from oracle import build_ref
o=Oracle('f32'); o.set_threads(1)
z=np.load('./tests/golden/c0_bg.npz')
case=Hh.case_from_meta(json.loads(str(z['meta'])), o)
got=Hh.run_operator(op, case, backward=False)
ref=Hh.run_operator(build_ref.import_reference(), case, backward=False)
orc=Hh.run_oracle(o, case, backward=False)
for nm,other in (('golden',z['radii']),('ref-live',ref['radii']),('oracle',orc['radii'])):
    bad=np.nonzero(got['radii']!=other)[0]
    print(nm, len(bad), [(int(i), int(got['radii'][i]), int(other[i])) for i in bad[:10]])
print('color maxdiff vs ref', np.abs(got['color']-ref['color']).max(), 'allmap', np.abs(got['allmap']-ref['allmap']).max())
