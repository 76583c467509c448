"""Key files (reference internal/blsgen/lib.go): the codec of the Python mirror (harmony_b200/blsgen.py) and of the C++
mirror (harmony_b200/host/hbls_keyfile.hpp, through hbls_host_cputest) against the reference's own vectors
(internal/blsgen/utils_test.go:30-43) and ten of its .hmy/*.key files (tests/golden/ref_fixtures.json "keyfiles").
CPU: decrypt -> sk, and the oracle's sk -> pk equals the file name.  GPU: the loaded key's GetPublicKey through the C ABI does."""
import json, os
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")))

def test_keyfile_codec_against_reference_files(oracle):
    from harmony_b200 import blsgen
    assert len(FIX["keyfiles"]) >= 10
    by_pk = {v["pk"]: v["sk"] for v in FIX["sk_pk"]}
    for kf in FIX["keyfiles"]:
        sk_hex = blsgen.decrypt(kf["blob"].encode(), kf["pass"]).decode()
        assert sk_hex == by_pk[kf["pk"]], kf["src"]
        assert oracle.get_public_key(bytes.fromhex(sk_hex)).hex() == kf["pk"]             # file name = hex(pk)
        raw = bytes.fromhex(kf["blob"])
        assert blsgen.encrypt(sk_hex.encode(), kf["pass"], nonce=raw[:12]) == kf["blob"]  # same nonce -> the reference's bytes
        assert blsgen.decrypt(raw, kf["pass"]).decode() == sk_hex                          # binary form fall-back (lib.go:129-136)
        with pytest.raises(ValueError, match="message authentication failed"):
            blsgen.decrypt(kf["blob"].encode(), kf["pass"] + "x")
    with pytest.raises(ValueError, match="the data is empty"): blsgen.decryptRaw(b"", "")

@pytest.mark.gpu
def test_load_key_file_and_derive_public_key(gbls, tmp_path):
    from harmony_b200 import blsgen
    for kf in FIX["keyfiles"][:4]:
        path = tmp_path / (kf["pk"] + ".key"); path.write_text(kf["blob"])
        sk = blsgen.LoadBLSKeyWithPassPhrase(str(path), " " + kf["pass"] + "\n")
        assert sk.GetPublicKey().SerializeToHexStr() == kf["pk"]
    sk, fn = blsgen.GenBLSKeyWithPassPhrase("pw", str(tmp_path))
    assert os.path.basename(fn) == sk.GetPublicKey().SerializeToHexStr() + ".key"
    assert blsgen.LoadBLSKeyWithPassPhrase(fn, "pw").IsEqual(sk)
    with pytest.raises(ValueError): blsgen.LoadBLSKeyWithPassPhrase(fn, "wrong")
