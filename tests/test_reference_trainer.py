"""CPU, authoring container only: the reference's UNMODIFIED trainer pipeline (src/trainer.py + optimizer.py + scheduler.py +
utils/* + configs/dtu/scan24.yml) runs on this repo's model through the import shims of dbw_b200.compat.  See
tests/ref_trainer_driver.py for what is swapped and why the render entry points are dummies when there is no GPU."""
import json
import os
import subprocess
import sys

import pytest

from tests._refextract import have_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_reference(), reason='needs the reference checkout (/root/reference)')
def test_unmodified_reference_trainer_runs_on_the_drop_in_model(tmp_path):
    run_dir = tmp_path / 'runs' / 'synthetic' / 'b200'
    run_dir.parent.mkdir(parents=True)
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_trainer_driver.py'), str(run_dir), '--iters', '3'],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('DRIVER_RESULT ')][-1]
    out = json.loads(line[len('DRIVER_RESULT '):])
    # the loss columns of configs/dtu/scan24.yml, in the reference's order (dbw.py:145-157)
    assert out['loss_names'] == ['loss_rgb', 'loss_perceptual', 'loss_parsimony', 'loss_tv', 'loss_overlap', 'loss_total']
    # optimizer.py:9-14: the `texture*` parameters form the second Adam group with their own learning rate
    assert out['param_groups'] == [7, 3] and out['lrs'] == [5e-3, 5e-2]
    assert {'textures', 'texture_bkg', 'texture_ground', 'S', 'T', 'R_6d', 'alpha_logit'} <= set(out['moved'])
    assert out['cur_epoch'] == 1 and out['model_name'] == 'dbw'
    # checkpoint keys = the reference's parameter / buffer names
    assert {'sq_eps', 'R_6d_ground', 'T_ground', 'S', 'R_6d', 'T', 'alpha_logit', 'texture_bkg', 'texture_ground', 'textures',
            'R_world', 'T_world', 'bkg_verts_uvs', 'ground_verts_uvs', 'sq_eta', 'sq_omega', 'block_faces_uvs',
            'block_verts_uvs'} <= set(out['checkpoint_keys'])
    assert out['train_metrics_rows'] >= 3 and out['val_metrics_rows'] == 2
    assert out['resumed_epoch_start'] == 2 and out['resume_equal']
    assert 'pytorch3d' in out['shimmed'] and 'toolz' in out['shimmed']


def test_compat_shims_are_thin_and_fail_loudly():
    """in a subprocess (the stand-ins must not leak into this session): real helpers work, placeholders raise when called"""
    code = r'''
import sys; sys.path.insert(0, %r)
import torch, dbw_b200, dbw_b200.compat as compat
names = compat.install()
import toolz
assert toolz.merge({'a': 1}, {'b': 2}) == {'a': 1, 'b': 2} and toolz.valmap(lambda v: v + 1, {'a': 1}) == {'a': 2}
assert toolz.keyfilter(lambda k: k != 'a', {'a': 1, 'b': 2}) == {'b': 2} and toolz.valfilter(lambda v: v > 1, {'a': 1, 'b': 2}) == {'b': 2}
if 'pytorch3d' in names:
    from pytorch3d.structures import Meshes
    from pytorch3d.utils import ico_sphere
    from pytorch3d.ops import SubdivideMeshes
    from pytorch3d.transforms import rotation_6d_to_matrix, matrix_to_rotation_6d, random_rotations
    from pytorch3d.io import save_ply
    m = ico_sphere(1)
    v, f = m.get_mesh_verts_faces(0)
    assert v.shape == (42, 3) and f.shape == (80, 3)
    v2, f2 = SubdivideMeshes()(m).get_mesh_verts_faces(0)
    assert v2.shape == (162, 3) and f2.shape == (320, 3)
    R = random_rotations(4)
    assert torch.allclose(rotation_6d_to_matrix(matrix_to_rotation_6d(R)), R, atol=1e-5)
    try:
        save_ply('x.ply', v)
        raise SystemExit('placeholder did not raise')
    except NotImplementedError:
        pass
if 'seaborn' in names:
    import seaborn as sns
    from matplotlib import colors as mplcolors
    from dbw_b200 import geometry as G
    pal = sns.color_palette('hls', 21)
    cm = mplcolors.LinearSegmentedColormap.from_list('Custom', [mplcolors.to_rgb('gold')] + pal[3:] + pal[:2])
    import numpy as np
    x = np.linspace(0, 1, 11)
    assert np.allclose(cm(x)[:, :3], G.fancy_cmap()(x))
print('OK')
''' % ROOT
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and 'OK' in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
