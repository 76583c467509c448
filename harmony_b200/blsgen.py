"""harmony_b200/blsgen.py -- Python mirror of the reference's BLS key files (internal/blsgen/lib.go:20-159).

A key file is  hex(nonce12 || AES-256-GCM(key, nonce12, hex(sk)))  with key = the 32 ASCII characters of hex(md5(passphrase)); its
name is hex(pk) + ".key".  Host-only codec (the C++ mirror is harmony_b200/host/hbls_keyfile.hpp); the secret key it yields goes
through the C ABI (SecretKey.DeserializeHexStr / GetPublicKey run on the GPU).  Same names and error strings as the reference."""
import hashlib, os
from cryptography.hazmat.primitives.ciphers.aead import AESGCM
from . import bls

def createHash(key: str) -> str:                      # lib.go:101-105
    return hashlib.md5(key.encode()).hexdigest()

def encrypt(data: bytes, passphrase: str, nonce: bytes = None) -> str:          # lib.go:107-118
    nonce = os.urandom(12) if nonce is None else nonce
    return (nonce + AESGCM(createHash(passphrase).encode()).encrypt(nonce, data, None)).hex()

def decryptRaw(data: bytes, passphrase: str) -> bytes:                           # lib.go:139-159
    if len(data) == 0: raise ValueError("unable to decrypt raw data with the provided passphrase; the data is empty")
    if len(data) < 12: raise ValueError("failed to decrypt raw data with the provided passphrase; the data size is invalid")
    try: return AESGCM(createHash(passphrase).encode()).decrypt(data[:12], data[12:], None)
    except Exception: raise ValueError("cipher: message authentication failed")

def decrypt(encrypted: bytes, passphrase: str) -> bytes:                         # lib.go:120-137: hex form, then the raw binary form
    try:
        return decryptRaw(bytes.fromhex(encrypted.decode("ascii")), passphrase)
    except (ValueError, UnicodeDecodeError) as e:
        err = e
    try: return decryptRaw(encrypted, passphrase)
    except ValueError: raise err

def LoadBLSKeyWithPassPhrase(fileName: str, passphrase: str) -> "bls.SecretKey":  # lib.go:51-71
    try: blob = open(fileName, "rb").read()
    except OSError as e: raise OSError(f"attempted to load from {fileName}: {e}")
    plain = decrypt(blob, passphrase.strip())
    sk = bls.SecretKey()
    try: sk.DeserializeHexStr(plain.decode())
    except Exception as e: raise ValueError(f"could not deserialize byte content of {fileName} as BLS secret key: {e}")
    return sk

def GenBLSKeyWithPassPhrase(passphrase: str, directory: str = "."):              # lib.go:20-34
    sk = bls.SecretKey(); sk.SetByCSPRNG()
    fileName = os.path.join(directory, sk.GetPublicKey().SerializeToHexStr() + ".key")
    with open(fileName, "w") as f: f.write(encrypt(sk.SerializeToHexStr().encode(), passphrase))
    return sk, fileName
