"""harmony_b200/slash.py -- the cryptographic part of double-sign slash verification, batched (SURVEY.md 8a R8 call site
staking/slash/double-sign.go:230-262; the ballot checks of :139-168 that need no chain state come with it).

Host mirror of slash.Verify for a queue of records: per record the reference deserialises each ballot's signature, adds up the
ballot's signer keys and calls VerifyHash(sum, ConstructCommitPayload(staking era, ballot.BlockHeaderHash, Height, ViewID)) -- two
pairing checks per record, one cgo call at a time.  Here the 2 R checks of R records are ONE hbls_verify_batch_status call; the key
sums use the library's G1 operations (BytesToBLSPublicKey's LRU + PublicKey.Add, as the reference does).

Out of scope (chain state, stays in Go): validator wrapper / banned status, epoch checks, committee lookup and the offender-address
match (double-sign.go:123-137,169-214), RLP hashing of the ballots, slashing economics.
"""
from dataclasses import dataclass, field
from typing import List, Optional
from . import bls
from .consensus import DeviceBackend as _ConsensusBackend

errSignerKeyNotRightSize = "bls keys from slash candidate not right side"          # double-sign.go:90
errSlashBlockNoConflict = "cannot slash for signing on non-conflicting blocks"      # :93
errNoMatchingDoubleSignKeys = "no matching double sign keys"                        # :87
errBallotsNotDiff = "ballots submitted must be different"                           # :271
errFailVerifySlash = "could not verify bls key signature on slash"                  # :270
errSigDeserialize = "err blsSignatureDeserialize"
errKeyDeserialize = "err blsPublicKeyDeserialize"

@dataclass
class Vote:                                  # double-sign.go:45-50
    SignerPubKeys: List[bytes]
    BlockHeaderHash: bytes
    Signature: bytes

@dataclass
class Evidence:                              # Moment + ConflictingVotes (double-sign.go:27-43)
    Epoch: int
    ShardID: int
    Height: int
    ViewID: int
    FirstVote: Vote
    SecondVote: Vote
    Offender: bytes = b""

@dataclass
class Record:                                # double-sign.go:52-57
    Evidence: Evidence
    Reporter: bytes = b""

class DeviceBackend(_ConsensusBackend):
    """consensus.DeviceBackend (verify_status; aggregate_keys = the sum of a ballot's signer keys, double-sign.go:241-249) + a decode
    probe used only to order two errors of one ballot."""
    def sig_decodes(self, sig: bytes) -> bool:
        try: bls.Sign().Deserialize(sig); return True
        except ValueError: return False

def VerifyBallots(records: List[Record], backend=None) -> List[Optional[str]]:
    """Per record None or the first error slash.Verify would return from its ballot checks (double-sign.go:139-168, 215-262), in
    the reference's order; the signature checks of all records run in one device call."""
    be = backend or DeviceBackend()
    out: List[Optional[str]] = [None] * len(records)
    pks, sigs, msgs, owner, pre = [], [], [], [], {}
    for r, rec in enumerate(records):
        ev = rec.Evidence; first, second = ev.FirstVote, ev.SecondVote
        if any(len(k) != bls.PublicKeySizeInBytes for k in list(first.SignerPubKeys) + list(second.SignerPubKeys)):
            out[r] = errSignerKeyNotRightSize; continue
        if bytes(first.BlockHeaderHash) == bytes(second.BlockHeaderHash): out[r] = errSlashBlockNoConflict; continue
        if not any(bytes(a) == bytes(b) for a in first.SignerPubKeys for b in second.SignerPubKeys):
            out[r] = errNoMatchingDoubleSignKeys; continue
        if (first.SignerPubKeys, bytes(first.BlockHeaderHash), bytes(first.Signature)) == (second.SignerPubKeys, bytes(second.BlockHeaderHash), bytes(second.Signature)):
            out[r] = errBallotsNotDiff; continue
        for b, ballot in enumerate((first, second)):
            # slash verification only happens in the staking era: 48-byte commit payload (double-sign.go:250-252)
            payload = bls.ConstructCommitPayload(True, bytes(ballot.BlockHeaderHash), ev.Height, ev.ViewID)
            try: apk = be.aggregate_keys(ballot.SignerPubKeys)
            except ValueError:
                pre[(r, b)] = errKeyDeserialize if (len(ballot.Signature) == 96 and be.sig_decodes(ballot.Signature)) else errSigDeserialize
                apk = bytes(48)
            pks.append(apk); sigs.append(bytes(ballot.Signature) if len(ballot.Signature) == 96 else bytes(96)); msgs.append(payload); owner.append((r, b))
            if len(ballot.Signature) != 96: pre.setdefault((r, b), errSigDeserialize)
    st = be.verify_status(pks, sigs, msgs) if pks else b""
    for k, (r, b) in enumerate(owner):
        if out[r] is not None: continue                          # the first ballot already failed
        if (r, b) in pre: out[r] = pre[(r, b)]
        elif st[k] == bls.VB_BAD_SIG_ENCODING: out[r] = errSigDeserialize
        elif st[k] != bls.VB_OK: out[r] = errFailVerifySlash
    return out
