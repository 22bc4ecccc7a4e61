"""Host logic: TF-free model loading and the drop-in classes' non-compute behaviour."""
import os
import pickle

import numpy as np
import pytest

import cosmopower_b200 as cp
from cosmopower_b200 import _loader, standins


def test_real_models_load_from_npz_and_pkl_agree():
    for name, (kind, path) in standins.REAL_MODELS.items():
        fields = _loader.NN_FIELDS if kind == "cosmopower_NN" else _loader.PCANN_FIELDS
        a = _loader.load_npz(path + ".npz", fields)
        assert len(a["W_"]) == a["n_layers"]
        if os.path.isfile(path + ".pkl"):
            b = _loader.attributes_from_list(_loader.load_pickle_list(path + ".pkl"), fields)
            assert list(a["parameters"]) == list(b["parameters"])
            for k in ("W_", "b_", "alphas_", "betas_"):
                for x, y in zip(a[k], b[k]):
                    assert np.array_equal(x, y)
            assert np.array_equal(a["features_mean_"], b["features_mean_"])
            assert isinstance(b["W_"], list)       # the ListWrapper global became a plain list


def test_standin_pickles_use_reference_layout():
    standins.build_standins()
    kind, path = standins.STANDIN_MODELS["cmb_TT_NN"]
    raw = open(path + ".pkl", "rb").read()
    assert b"tensorflow.python.training.tracking.data_structures" in raw[:200]
    m = cp.cosmopower_NN(restore=True, restore_filename=path)
    assert m.architecture == [6, 512, 512, 512, 512, 2507]
    assert m.n_layers == 5 and m.n_modes == 2507 and len(m.modes) == 2507
    assert [w.dtype for w in m.W_] == [np.float32] * 5
    te = cp.cosmopower_PCAplusNN(restore=True, restore_filename=standins.STANDIN_MODELS["cmb_TE_PCAplusNN"][1])
    assert te.n_pcas == 512 and te.pca_transform_matrix_.shape == (512, 2507)


def test_attribute_surface_matches_reference():
    kind, path = standins.model_path("cmb_PP_PCAplusNN")
    m = cp.cosmopower_PCAplusNN(restore=True, restore_filename=path)
    for k in ("parameters", "n_parameters", "modes", "n_modes", "n_hidden", "n_layers", "architecture", "W_",
              "b_", "alphas_", "betas_", "parameters_mean_", "parameters_std_", "features_mean_",
              "features_std_", "n_pcas", "pca_mean_", "pca_std_", "pca_transform_matrix_"):
        assert hasattr(m, k), k
    assert m.parameters == ["omega_b", "omega_cdm", "h", "tau_reio", "n_s", "ln10^{10}A_s"]
    assert m.n_pcas == 64


def test_wrong_class_raises_value_error_like_reference():
    _, pp = standins.model_path("cmb_PP_PCAplusNN")
    _, pk = standins.model_path("PKLIN_NN")
    with pytest.raises(ValueError):
        cp.cosmopower_NN(restore=True, restore_filename=pp)
    with pytest.raises(ValueError):
        cp.cosmopower_PCAplusNN(restore=True, restore_filename=pk)


def test_missing_file_raises_file_not_found():
    with pytest.raises(FileNotFoundError):
        cp.cosmopower_NN(restore=True, restore_filename="/nonexistent/model")


def test_unpickler_refuses_foreign_globals(tmp_path):
    p = tmp_path / "evil.pkl"
    with open(p, "wb") as f:
        pickle.dump([os.system], f)
    with pytest.raises(pickle.UnpicklingError):
        _loader.load_pickle_list(str(p))


def test_dict_to_ordered_arr_np_semantics():
    _, pk = standins.model_path("PKLIN_NN")
    m = cp.cosmopower_NN(restore=True, restore_filename=pk)
    d = {k: [float(i), float(i) + 0.5] for i, k in enumerate(m.parameters)}
    d["unused"] = [9., 9.]
    arr = m.dict_to_ordered_arr_np(d)
    assert arr.shape == (2, 6) and arr.dtype == np.float64
    assert np.array_equal(arr[0], np.arange(6.))
    d32 = {k: np.array(v, dtype=np.float32) for k, v in d.items()}
    assert m.dict_to_ordered_arr_np(d32).dtype == np.float32
    with pytest.raises(KeyError):
        m.dict_to_ordered_arr_np({"omega_b": [1.0]})


def test_npz_roundtrip(tmp_path):
    _, pk = standins.model_path("PKNLBOOST_NN")
    m = cp.cosmopower_NN(restore=True, restore_filename=pk)
    m.save(str(tmp_path / "copy"))
    m2 = cp.cosmopower_NN(restore=True, restore_filename=str(tmp_path / "copy"))
    assert m2.parameters == m.parameters
    for a, b in zip(m.W_, m2.W_):
        assert np.array_equal(a, b)


def test_non_restore_constructor_with_weights():
    rng = np.random.default_rng(0)
    w = dict(W_=[rng.standard_normal((3, 8)), rng.standard_normal((8, 5))],
             b_=[np.zeros(8), np.zeros(5)], alphas_=[np.ones(8)], betas_=[np.zeros(8)])
    m = cp.cosmopower_NN(parameters=["a", "b", "c"], modes=np.arange(5), n_hidden=[8], weights=w)
    assert m.architecture == [3, 8, 5] and m.n_layers == 2
    with pytest.raises(NotImplementedError):
        cp.cosmopower_NN(parameters=["a"], modes=np.arange(5), n_hidden=[8])
    with pytest.raises(NotImplementedError):
        m.train()


def test_rescaled_predictions_follow_the_reference_contract():
    """reference cosmopower_NN.py:429-466: `rescaled_predictions_np` applies `postprocessing_np` (set by train(), never pickled)
    to the network output; a restored model has no such attribute and raises AttributeError, here as there."""
    kind, path = standins.REAL_MODELS["PKLIN_NN"]
    m = cp.cosmopower_NN(restore=True, restore_filename=path)
    for call in (m.rescaled_predictions_np, m.ten_to_rescaled_predictions_np):
        with pytest.raises(AttributeError):
            call({k: np.zeros(1) for k in m.parameters})
    with pytest.raises(AttributeError):
        m.processing_vectors_tf
    n = int(m.n_modes)
    y = np.linspace(-1., 1., 3 * n).reshape(3, n)
    mean, sigma = np.arange(n) * 0.5, 1. + np.arange(n) * 0.01
    m.set_feature_processing('mean_sigma', {'mean': mean, 'sigma': sigma})       # train(): cosmopower_NN.py:666-677
    assert np.array_equal(m.postprocessing_np(y, m.processing_vectors_np), y * sigma + mean)
    lo, hi = -np.ones(n), 2. + np.arange(n) * 0.1
    m.set_feature_processing('min_max', {'min': lo, 'max': hi})                   # :680-691
    assert np.array_equal(m.postprocessing_np(y, m.processing_vectors_np), y * (hi - lo) + lo)
    m.set_feature_processing(None)                                               # :694-703
    assert m.postprocessing_np(y, m.processing_vectors_np) is y and m.processing_vectors_np == {}
    with pytest.raises(ValueError):
        m.set_feature_processing('mean_sigma', {'mean': mean[:-1], 'sigma': sigma[:-1]})
    with pytest.raises(ValueError):
        m.set_feature_processing('standard')


def test_activation_helper_matches_the_reference_formula():
    kind, path = standins.REAL_MODELS["PKLIN_NN"]
    m = cp.cosmopower_NN(restore=True, restore_filename=path)
    x = np.linspace(-30., 30., 13)
    a, b = 0.7, 0.2
    want = (b + (1. / (1. + np.exp(-a * x))) * (1. - b)) * x          # cosmopower_NN.py:156
    np.testing.assert_allclose(m.activation(x, a, b), want, rtol=1e-15)
    assert np.all(np.isfinite(m.activation(np.array([-1e4, 1e4]), 1.0, 0.5)))
    assert m.update_emulator_parameters() is None
