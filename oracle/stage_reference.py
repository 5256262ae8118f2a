"""Stage the UNMODIFIED reference path under oracle/_ref/.  TEST INFRASTRUCTURE.

    python -m oracle.stage_reference

The reference (NVlabs/bayes-sim-ig) is a pure-Python package, so "building" its
implementation of the hot path is a verbatim file copy of the modules the path
consists of (bayes_sim.py, models/{mdnn,mdrff,rff}.py, utils/{summarizers,pdf}.py) from
/root/reference into oracle/_ref/bayes_sim_ig/.  oracle/_ref/ is git-ignored (reference
sources never enter this repository's history) but travels to the GPU box with the
snapshot, exactly like the built libbsig_b200.so, so that ``bench.py --impl reference``
and the ``cpu_baseline`` leg time the reference's own code (``kind: "reference"``)
instead of the torch port.  ``__graft_entry__.build()`` runs this whenever
/root/reference is present; nothing is modified on the way (checked by SHA-256).
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = '/root/reference'
DST_ROOT = os.path.join(HERE, '_ref')
FILES = ['bayes_sim.py', 'models/mdnn.py', 'models/mdrff.py', 'models/rff.py',
         'utils/summarizers.py', 'utils/pdf.py']


def _sha(path):
    with open(path, 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(src_root=SRC_ROOT, dst_root=DST_ROOT):
    """Copy the path's modules; returns {relative path: sha256} or None when the
    reference is not mounted (GPU box: the staged copy that travelled is used)."""
    src_pkg = os.path.join(src_root, 'bayes_sim_ig')
    if not os.path.isdir(src_pkg):
        return None
    sums = {}
    for rel in FILES:
        src = os.path.join(src_pkg, rel)
        dst = os.path.join(dst_root, 'bayes_sim_ig', rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        assert _sha(src) == _sha(dst)
        sums[rel] = _sha(dst)
    with open(os.path.join(dst_root, 'MANIFEST.txt'), 'w') as fh:
        fh.write('verbatim copies of %s/bayes_sim_ig (sha256)\n' % src_root)
        for rel in FILES:
            fh.write('%s  %s\n' % (sums[rel], rel))
    return sums


def staged():
    return all(os.path.exists(os.path.join(DST_ROOT, 'bayes_sim_ig', rel)) for rel in FILES)


if __name__ == '__main__':
    print(stage() or 'reference not mounted; staged copy present: %s' % staged())
