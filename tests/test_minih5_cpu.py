"""minih5: the minimal HDF5 writer behind prediction_io.save_batch_nn (h5py is not part of the image).
Pins: the lookup3 checksum against the known answers printed by Bob Jenkins' lookup3.c (driver2), the byte layout of the
superblock / object headers against the HDF5 file-format specification (field by field), exact write -> read round trips."""
import struct

import numpy as np
import pytest

from articulated_pose_b200 import minih5, prediction_io


def test_lookup3_known_answers():
    s = b"Four score and seven years ago"
    assert minih5.lookup3(s, 0) == 0x17770551 and minih5.lookup3(s, 1) == 0xCD628161
    assert minih5.lookup3(b"", 0) == 0xDEADBEEF and minih5.lookup3(b"", 0xDEADBEEF) == 0xBD5B7DDE


def test_file_layout_follows_the_specification(tmp_path):
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    p = str(tmp_path / "x.h5")
    minih5.write(p, {"P": a, "cls": np.array([1, 2, 3], np.int64)}, {"basename": "0001_0_0"})
    buf = open(p, "rb").read()
    # superblock v2: signature, version 2, 8-byte offsets and lengths, base 0, no extension, eof, root address, checksum
    assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8:12] == bytes([2, 8, 8, 0])
    base, ext, eof, root = struct.unpack_from("<QQQQ", buf, 12)
    assert (base, ext, eof, root) == (0, 0xFFFFFFFFFFFFFFFF, len(buf), 48)
    assert struct.unpack_from("<I", buf, 44)[0] == minih5.lookup3(buf[:44])
    # root object header: 'OHDR', version 2, 4-byte chunk size, checksum over prefix + messages
    assert buf[48:52] == b"OHDR" and buf[52] == 2 and buf[53] == 0x02
    size = struct.unpack_from("<I", buf, 54)[0]
    assert struct.unpack_from("<I", buf, 58 + size)[0] == minih5.lookup3(buf[48:58 + size])
    msgs = minih5._parse_header(buf, 48)
    assert [m[0] for m in msgs] == [0x02, 0x0A, 0x06, 0x06, 0x0C]          # link info, group info, 2 links, 1 attribute
    assert msgs[0][1] == bytes([0, 0]) + b"\xff" * 16                      # compact link storage: no heap, no index
    link = msgs[2][1]
    assert link[:4] == bytes([1, 0x10, 1, 1]) and link[4:5] == b"P"        # v1, charset field, UTF-8, name length 1
    ds_addr = struct.unpack_from("<Q", link, 5)[0]
    dm = dict(minih5._parse_header(buf, ds_addr))
    assert dm[0x01] == bytes([2, 3, 0, 1]) + struct.pack("<QQQ", 2, 3, 4)  # dataspace v2, rank 3, simple
    assert dm[0x03] == bytes([0x11, 0x20, 31, 0]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)   # IEEE f32 LE
    assert dm[0x05] == bytes([3, 0x0A])
    ver, cls, addr, nbytes = struct.unpack_from("<BBQQ", dm[0x08], 0)
    assert (ver, cls, nbytes) == (3, 1, a.nbytes) and addr % 8 == 0
    assert np.array_equal(np.frombuffer(buf, "<f4", a.size, addr).reshape(a.shape), a)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.int64, np.uint8, np.int16, np.bool_])
def test_roundtrip_dtypes(tmp_path, dt):
    rng = np.random.default_rng(0)
    a = (rng.normal(size=(7, 5)) * 50).astype(dt)
    p = str(tmp_path / "r.h5")
    minih5.write(p, {"a": a, "scalar": np.asarray(a.flat[0]), "empty": np.zeros((0, 3), dt)}, {"method_name": "ancsh"})
    d, attrs = minih5.read(p)
    exp = a.astype(np.uint8) if dt is np.bool_ else a
    assert d["a"].dtype == exp.dtype and np.array_equal(d["a"], exp)
    assert d["scalar"].shape == () and d["empty"].shape == (0, 3) and attrs == {"method_name": "ancsh"}


def test_corruption_is_detected(tmp_path):
    p = str(tmp_path / "c.h5")
    minih5.write(p, {"a": np.arange(4.0)})
    buf = bytearray(open(p, "rb").read())
    buf[60] ^= 0x40
    open(p, "wb").write(buf)
    with pytest.raises(minih5.FormatError):
        minih5.read(p)


def test_save_batch_nn_writes_h5_files_the_loader_reads(tmp_path):
    B, N, K = 2, 16, 3
    rng = np.random.default_rng(1)
    pred = {"W": rng.random((B, N, K), dtype=np.float32), "confi_per_point": rng.random((B, N, 1), dtype=np.float32),
            "nocs_per_point": rng.random((B, N, 3 * K), dtype=np.float32), "gocs_per_point": rng.random((B, N, 3 * K), dtype=np.float32),
            "heatmap_per_point": rng.random((B, N, 1), dtype=np.float32), "unitvec_per_point": rng.random((B, N, 3), dtype=np.float32),
            "joint_axis_per_point": rng.random((B, N, 3), dtype=np.float32), "index_per_point": rng.random((B, N, 3), dtype=np.float32)}
    inp = {"P": rng.random((B, N, 3), dtype=np.float32), "cls_gt": rng.integers(0, K, (B, N)).astype(np.float32),
           "nocs_gt": rng.random((B, N, 3), dtype=np.float32), "nocs_gt_g": rng.random((B, N, 3), dtype=np.float32),
           "heatmap_gt": rng.random((B, N), dtype=np.float32), "unitvec_gt": rng.random((B, N, 3), dtype=np.float32),
           "orient_gt": rng.random((B, N, 3), dtype=np.float32), "joint_cls_gt": rng.integers(0, K, (B, N)).astype(np.float32)}
    names = ["0001_0_0", "0001_0_1"]
    prediction_io.save_batch_nn("ancsh", pred, inp, names, str(tmp_path), is_mixed=True, W_reduced=True)
    assert prediction_io.list_predictions(str(tmp_path)) == [n + ".h5" for n in names]
    assert open(str(tmp_path / "0001_0_0.h5"), "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    f = prediction_io.load_prediction(str(tmp_path), names[1])
    assert set(f.keys()) == set(prediction_io.DATASETS)
    np.testing.assert_array_equal(f["nocs_per_point"][()], pred["nocs_per_point"][1])
    np.testing.assert_array_equal(f["instance_per_point"][()], np.argmax(pred["W"][1], 1))       # int64 labels (W_reduced)
    np.testing.assert_array_equal(f["P"][3:5, :3], inp["P"][1, 3:5])
    assert f.attrs["basename"] == names[1] and f.attrs["method_name"] == "ancsh"
