#!/usr/bin/env python3
"""Write a minimal synthetic tekken.json next to the synthetic checkpoint.

The reference tokenizer (voxtral_tokenizer.c:5-13 in /root/reference) only reads
`vocab[].rank`, `vocab[].token_bytes` (base64) and `special_tokens[].rank/.token_str`.
IDs 0..999 are special tokens, IDs 1000.. map to vocab[id-1000].

Rank 0 is the single byte 0x00 like the real Tekken vocabulary, which the
stream code classifies as an *invalid* (empty) text token
(voxtral.c:489-497) -- the synthetic vocabulary keeps that edge case.
"""
import base64
import json
import sys

N_VOCAB = 130072
N_SPECIAL = 1000


def piece(rank: int) -> bytes:
    if rank < 256:
        return bytes([rank])
    # deterministic pseudo words; every 7th piece starts a new word with a space
    word = "w%x" % rank
    return ((" " if rank % 7 == 0 else "") + word).encode()


def main(out_dir: str) -> None:
    vocab = [
        {"rank": r, "token_bytes": base64.b64encode(piece(r)).decode()}
        for r in range(N_VOCAB)
    ]
    names = {0: "<unk>", 1: "<s>", 2: "</s>", 32: "[STREAMING_PAD]", 33: "[STREAMING_WORD]"}
    special = [
        {"rank": r, "token_str": names.get(r, "<SPECIAL_%d>" % r), "is_control": True}
        for r in range(N_SPECIAL)
    ]
    doc = {
        "config": {"default_vocab_size": N_VOCAB + N_SPECIAL, "default_num_special_tokens": N_SPECIAL},
        "vocab": vocab,
        "special_tokens": special,
    }
    with open(out_dir.rstrip("/") + "/tekken.json", "w") as f:
        json.dump(doc, f, separators=(",", ":"))


if __name__ == "__main__":
    main(sys.argv[1])
