"""CPU-only checks: the C-ABI library exports every symbol include/glamr_b200.h declares (no compute calls), struct
layouts agree with the ctypes mirror, the built-in stage tables equal the reference YAML, the product refuses CPU."""
import ctypes
import numpy as np
import os

import pytest
import torch
import yaml

import __graft_entry__ as ge
from conftest import REFERENCE_ROOT
from glamr_b200 import lib as L
from glamr_b200.config import BUILTIN_IDS, Config, builtin_config_dict


@pytest.fixture(scope='module')
def cdll():
    ge.build()
    return ctypes.CDLL(L.SO_PATH)


def test_library_exports_every_declared_symbol(cdll):
    syms = ge.exported_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(cdll, s), s


def test_struct_layouts_match(cdll):
    cdll.glamr_sizeof_person.restype = ctypes.c_size_t
    cdll.glamr_sizeof_problem.restype = ctypes.c_size_t
    assert cdll.glamr_sizeof_person() == ctypes.sizeof(L.Person)
    assert cdll.glamr_sizeof_problem() == ctypes.sizeof(L.Problem)
    assert cdll.glamr_version() >= 100


def test_term_table_matches_header():
    hdr = open(os.path.join(os.path.dirname(L.HERE), 'include', 'glamr_b200.h')).read()
    body = hdr[hdr.index('enum glamr_term {'):hdr.index('GLAMR_NUM_TERMS')]
    names = [n.strip().split('=')[0].strip() for n in body.split('{')[1].split(',') if n.strip()]
    assert len(names) == L.NUM_TERMS == len(L.TERM_INDEX)


def test_product_requires_cuda_device():
    with pytest.raises(L.GlamrError):
        L.require_cuda('cpu')


@pytest.mark.reference
@pytest.mark.parametrize('cfg_id', BUILTIN_IDS)
def test_builtin_configs_equal_reference_yaml(cfg_id):
    ref = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, 'global_recon', 'cfg', cfg_id + '.yml')))
    mine = builtin_config_dict(cfg_id)
    assert mine['grecon_model_specs'] == ref['grecon_model_specs']
    assert mine['opt_stage_specs'] == ref['opt_stage_specs']
    assert mine['grecon_model_name'] == ref['grecon_model_name'] and mine['dataset'] == ref['dataset']


def test_config_surface():
    cfg = Config('glamr_static_multi', out_dir='/tmp/glamr_b200_cfg_test')
    assert cfg.id == 'glamr_static_multi' and cfg.grecon_model_name == 'global_recon_model'
    assert list(cfg.opt_stage_specs) == ['init_opt', 'main_opt'] and cfg.grecon_model_specs['flag_fixed_cam'] is True


def test_rotmats_to_rotvec_matches_scipy():
    """host step of init_data (global_recon_model.py:106-107): float32 rotation matrices -> rotation vectors"""
    from scipy.spatial.transform import Rotation
    from glamr_b200.recon import rotmats_to_rotvec
    rng = np.random.default_rng(0)
    rv = rng.normal(size=(4000, 3)) * rng.uniform(0, 1.5, size=(4000, 1))
    rv[:50] *= 1e-5                                                                       # series branch
    rv[50:100] = rv[50:100] / np.linalg.norm(rv[50:100], axis=1, keepdims=True) * (np.pi - 1e-4)   # near pi
    mats = Rotation.from_rotvec(rv).as_matrix().astype(np.float32)                        # float32-accurate, as HybrIK stores them
    np.testing.assert_allclose(rotmats_to_rotvec(mats), Rotation.from_matrix(mats).as_rotvec(), atol=1e-10)
