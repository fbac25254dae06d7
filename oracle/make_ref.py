"""Copy the UNMODIFIED reference implementation of the hot path into oracle/_ref/ (git-ignored, travels to the GPU box with
the snapshot) so that `bench.py --impl reference` can time the reference's own code on the box's host cores
(`cpu_baseline.kind = "reference"`).  Only the files the path imports are copied, byte for byte:

    config.py  utils.py  models/**  common/loss.py  common/weight_init.py

The reference is a pure-Python script tree (no setup.py, nothing to compile), so there is no build step; it is imported
with the recipe of SURVEY.md section 8c (cwd = the tree, sys.argv set before `import config`).  Runs only where
/root/reference exists (the build container); on the GPU box the prepared copy is used as is.

    python oracle/make_ref.py
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
DST = os.path.join(HERE, '_ref')
FILES = ['config.py', 'utils.py', 'common/loss.py', 'common/weight_init.py']
TREES = ['models']


def main() -> int:
    if not os.path.isdir(REF):
        print(f'{REF} not present: keeping {DST} as it is')
        return 0 if os.path.isdir(DST) else 1
    os.makedirs(DST, exist_ok=True)
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for rel in TREES:
        dst = os.path.join(DST, rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, rel), dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    # verify: byte-identical to the source tree
    bad = [rel for rel in FILES if not filecmp.cmp(os.path.join(REF, rel), os.path.join(DST, rel), shallow=False)]
    for rel in TREES:
        for root, _dirs, files in os.walk(os.path.join(REF, rel)):
            for f in files:
                if f.endswith('.pyc'):
                    continue
                src = os.path.join(root, f)
                if not filecmp.cmp(src, os.path.join(DST, os.path.relpath(src, REF)), shallow=False):
                    bad.append(os.path.relpath(src, REF))
    if bad:
        print('copy differs from the reference:', bad)
        return 1
    print(f'reference hot-path files copied unmodified to {DST}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
