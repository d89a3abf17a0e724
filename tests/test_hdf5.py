"""ursonet_b200/hdf5.py: the pure-Python HDF5 subset behind Keras `.h5` weight files (net.py:816-852, 1120).

Pin: tests/golden/testhdf5_7.4_GLNX86.mat is a file written by libhdf5 itself (MATLAB v7.3 format, copied from scipy's
test data, scipy/io/matlab/tests/data; expected contents per scipy's test_mio.py: 'testdouble' = pi/4 * arange(9) as a
1x9 MATLAB array, stored transposed).  The reader must read it (also in strict mode, which checks the invariants the
writer is held to), and the writer's encoders must reproduce that file's structures byte for byte.  No libhdf5 is
available to open the writer's files: "writer unverified against libhdf5"."""
import os
import struct

import numpy as np
import pytest

from ursonet_b200 import hdf5

GOLD = os.path.join(os.path.dirname(__file__), "golden", "testhdf5_7.4_GLNX86.mat")
BASE = 512          # the MATLAB user block: every address in the file is relative to it


def test_reader_on_a_libhdf5_file():
    for strict in (False, True):
        with hdf5.File(GOLD, strict=strict) as f:
            assert f.superblock["version"] == 0 and f.superblock["base"] == BASE
            assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
            d = f["testdouble"]
            assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
            assert d.attrs == {"MATLAB_class": b"double"}
            np.testing.assert_array_equal(d.read().ravel(), np.pi / 4 * np.arange(9))


def test_encoders_match_libhdf5_bytes():
    raw = open(GOLD, "rb").read()
    # dataset object header at 0x5d0: fill value (0x5e8), datatype (0x5f8), dataspace (0x618), attribute (0x670)
    assert raw[0x5f8:0x5f8 + 20] == hdf5.encode_datatype(np.float64)
    assert raw[0x618:0x618 + 24] == hdf5.encode_dataspace((9, 1))
    assert raw[0x670:0x670 + 46] == hdf5.encode_attribute("MATLAB_class", b"double", strpad=0)
    assert raw[0x5e8:0x5f0] == bytes([1, 2, 2, 1, 0, 0, 0, 0])          # the fill-value message the writer emits
    # root group object header at 0x5a0: prefix + symbol table message
    mine = hdf5.encode_object_header([(hdf5.MSG_SYMBOL_TABLE, 1, struct.pack("<QQ", 0x180, 0x60)), (hdf5.MSG_NIL, 0, b"")])
    assert raw[0x5a0:0x5a0 + 16 + 32] == mine        # libhdf5 pads the root header with one empty NIL message


def test_group_structures_match_libhdf5(tmp_path):
    """One group with the one child 'testdouble': local heap, B-tree node and symbol table node equal the libhdf5 file's
    up to addresses."""
    p = str(tmp_path / "t.h5")
    hdf5.write_file(p, {"testdouble": ("d", (np.pi / 4 * np.arange(9)).reshape(9, 1), {"MATLAB_class": b"double"})})
    mine, gold = open(p, "rb").read(), open(GOLD, "rb").read()
    with hdf5.File(p, strict=True) as f:
        np.testing.assert_array_equal(f["testdouble"].read().ravel(), np.pi / 4 * np.arange(9))
    # superblock: same bytes except consistency flags (a MATLAB quirk), addresses and the root entry
    assert mine[:20] == gold[BASE:BASE + 20]
    u = lambda b, o: int.from_bytes(b[o:o + 8], "little")
    m_bt, m_hp, g_bt, g_hp = u(mine, 80), u(mine, 88), BASE + u(gold, BASE + 80), BASE + u(gold, BASE + 88)
    assert mine[72:76] == gold[BASE + 72:BASE + 76] == struct.pack("<I", 1)        # cached symbol table in the root entry
    # local heap header (signature, version, free-list head) and the used part of its data segment
    assert mine[m_hp:m_hp + 8] == gold[g_hp:g_hp + 8]
    assert u(mine, m_hp + 16) == u(gold, g_hp + 16) == 24
    m_seg, g_seg = u(mine, m_hp + 24), BASE + u(gold, g_hp + 24)
    assert mine[m_seg:m_seg + 32] == gold[g_seg:g_seg + 32]                        # '', 'testdouble', free block 'next' = 1
    assert u(mine, m_seg + 32) == u(mine, m_hp + 8) - 24 and u(gold, g_seg + 32) == u(gold, g_hp + 8) - 24
    # B-tree node: header, siblings, key 0, (child), key 1
    assert mine[m_bt:m_bt + 32] == gold[g_bt:g_bt + 32]
    assert mine[m_bt + 40:m_bt + 48] == gold[g_bt + 40:g_bt + 48] == struct.pack("<Q", 8)
    # symbol table node: header and the entry (name offset, cache type, scratch) up to the object address
    m_sn, g_sn = u(mine, m_bt + 32), BASE + u(gold, g_bt + 32)
    assert mine[m_sn:m_sn + 16] == gold[g_sn:g_sn + 16]
    assert mine[m_sn + 24:m_sn + 48] == gold[g_sn + 24:g_sn + 48]
    assert len(mine) == u(mine, 40)                                                # end-of-file address


def _state(n_layers, rng):
    sd = {}
    for i in range(n_layers):
        sd["res%03d_branch2a/kernel" % i] = rng.standard_normal((3, 3, 4, 8)).astype(np.float32)
        sd["res%03d_branch2a/bias" % i] = rng.standard_normal(8).astype(np.float32)
        for k in ("gamma", "beta", "moving_mean", "moving_variance"):
            sd["bn%03d_branch2a/%s" % (i, k)] = rng.standard_normal(8).astype(np.float32)
    return sd


@pytest.mark.parametrize("n_layers", [0, 1, 4, 5, 40, 150, 700])
def test_keras_weight_file_round_trip(tmp_path, n_layers):
    """1 / 2 / many symbol-table nodes, and (700 layers = 1400 groups) a two-level B-tree."""
    sd = _state(n_layers, np.random.default_rng(n_layers))
    p = str(tmp_path / "w.h5")
    hdf5.write_keras_weights(p, sd)
    with hdf5.File(p, strict=True) as f:
        assert len(f.keys()) == 2 * n_layers
        assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.1.6"
        if n_layers:
            assert [bytes(x).decode() for x in f.attrs["layer_names"]][:2] == ["res000_branch2a", "bn000_branch2a"]
            g = f["bn000_branch2a"]
            assert [bytes(x) for x in g.attrs["weight_names"]] == [b"bn000_branch2a/gamma:0", b"bn000_branch2a/beta:0",
                                                                    b"bn000_branch2a/moving_mean:0",
                                                                    b"bn000_branch2a/moving_variance:0"]
            assert g["bn000_branch2a/gamma:0"].shape == (8,)
    back = hdf5.keras_to_state_dict(hdf5.read_keras_weights(p))
    assert list(back) == list(sd)
    for k in sd:
        np.testing.assert_array_equal(back[k], sd[k])


def test_old_keras_weight_names_map_by_position(tmp_path):
    """keras-applications' ResNet-50 file predates the '<layer>/kernel:0' names ('conv1_W_1:0', 'bn_conv1_running_std_1:0');
    Keras assigns by position in weight_names, so does the loader.  Also a full-model file ('model_weights' group)."""
    rng = np.random.default_rng(1)
    w, b = rng.standard_normal((7, 7, 3, 64)).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    bn = [rng.standard_normal(64).astype(np.float32) for _ in range(4)]
    conv = ("g", {"conv1_W_1:0": ("d", w, {}), "conv1_b_1:0": ("d", b, {})},
            {"weight_names": np.array([b"conv1_W_1:0", b"conv1_b_1:0"])})
    names = [b"bn_conv1_gamma_1:0", b"bn_conv1_beta_1:0", b"bn_conv1_running_mean_1:0", b"bn_conv1_running_std_1:0"]
    bng = ("g", {n.decode(): ("d", a, {}) for n, a in zip(names, bn)}, {"weight_names": np.array(names)})
    empty = ("g", {}, {"weight_names": np.zeros((0,), "S1")})
    layers = {"conv1": conv, "bn_conv1": bng, "activation_1": empty}
    # long layer lists are split into layer_names0, layer_names1, ... by Keras (HDF5's 64 KB header limit)
    attrs = {"layer_names0": np.array([b"conv1", b"activation_1"]), "layer_names1": np.array([b"bn_conv1"])}
    p = str(tmp_path / "model.h5")
    hdf5.write_file(p, {"model_weights": ("g", layers, attrs)}, {"keras_version": b"2.0.8"})
    sd = hdf5.keras_to_state_dict(hdf5.read_keras_weights(p))
    assert list(sd) == ["conv1/kernel", "conv1/bias", "bn_conv1/gamma", "bn_conv1/beta", "bn_conv1/moving_mean",
                        "bn_conv1/moving_variance"]
    np.testing.assert_array_equal(sd["conv1/kernel"], w)
    np.testing.assert_array_equal(sd["bn_conv1/moving_variance"], bn[3])


def test_dtypes_scalars_and_errors(tmp_path):
    p = str(tmp_path / "d.h5")
    tree = {"i": ("d", np.arange(6, dtype=np.int64).reshape(2, 3), {"n": np.int32(7), "v": np.arange(3, dtype=np.float64)}),
            "h": ("d", np.arange(4, dtype=np.float16), {}), "s": ("d", np.float32(2.5), {}),
            "u": ("d", np.arange(5, dtype=np.uint8), {}), "e": ("d", np.zeros((0, 4), np.float32), {}),
            "be": ("d", np.arange(3, dtype=">f4"), {})}
    hdf5.write_file(p, tree, {"note": "text"})
    with hdf5.File(p, strict=True) as f:
        assert f.attrs["note"] == b"text"
        np.testing.assert_array_equal(f["i"].read(), np.arange(6).reshape(2, 3))
        assert f["i"].attrs["n"] == 7 and f["i"].attrs["v"].tolist() == [0.0, 1.0, 2.0]
        assert f["h"].read().dtype == np.float16 and f["s"].read() == np.float32(2.5) and f["s"].shape == ()
        assert f["u"].read().tolist() == [0, 1, 2, 3, 4] and f["e"].read().shape == (0, 4)
        assert f["be"].read().tolist() == [0.0, 1.0, 2.0]
        with pytest.raises(KeyError):
            f["i/x"]
    raw = bytearray(open(p, "rb").read())
    for bad, msg in ((raw[:200], "corrupt|beyond|truncated"), (b"not an hdf5 file" * 100, "not an HDF5 file"),
                     (bytes(raw[:8]) + b"\x02" + bytes(raw[9:]), "superblock version 2")):
        q = str(tmp_path / "bad.h5")
        open(q, "wb").write(bytes(bad))
        with pytest.raises(hdf5.Hdf5Error, match=msg):
            hdf5.File(q)
    with pytest.raises(hdf5.Hdf5Error, match="64 KB"):
        hdf5.write_file(p, {}, {"big": np.zeros(70000, np.uint8)})
