"""CPU ORACLE (test infrastructure only) for the GRU text generator of
/root/reference/rnn_text_gen/rnn_text_generation.cpp -- numpy restatement of gru_forward (:186-263) and the greedy loop
of inference (:266-314).  PARITY UNPINNED (the reference ships no tests or weights; gru.bin comes from a TF training run).

File layout of gru.bin (rnn_text_generation.py:101-115, read back at rnn.cpp:117-147): for each of
embeddings (66,256), gru kernel (256,3072), recurrent kernel (1024,3072), bias (2,3072), dense kernel (1024,66):
int32 n_dims, int32 dims[2], f32 data;  dense bias (66,): int32 n_dims, int32 dim, f32 data."""
import struct

import numpy as np

VOCAB = "\t\n !$&',-.3:;?ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"  # rnn.cpp:22


def make_synthetic_gru(seed=5, vocab=66, emb=256, units=1024):
    rng = np.random.default_rng(seed)
    f = np.float32
    return {
        "emb": rng.normal(0, 1.0, (vocab, emb)).astype(f),
        "W": rng.normal(0, 1.0 / np.sqrt(emb), (emb, 3 * units)).astype(f),
        "U": rng.normal(0, 1.0 / np.sqrt(units), (units, 3 * units)).astype(f),
        "b": rng.normal(0, 0.1, (2, 3 * units)).astype(f),
        "D": rng.normal(0, 1.0 / np.sqrt(units), (units, vocab)).astype(f),
        "c": rng.normal(0, 0.1, (vocab,)).astype(f),
    }


def write_gru_bin(path, w):
    with open(path, "wb") as fh:
        for k in ("emb", "W", "U", "b", "D"):
            a = w[k]
            fh.write(struct.pack("iii", 2, a.shape[1], a.shape[0]))  # n_dims + dims (the reader skips these 3 ints)
            fh.write(a.tobytes())
        fh.write(struct.pack("ii", 1, w["c"].shape[0]))
        fh.write(w["c"].tobytes())


def _sigmoid_like_reference(x):
    # rnn.cpp:51-55: sigmoid(x) computed as silu(x) / x  (NaN at exactly 0; irrelevant for random weights)
    x = x.astype(np.float32)
    silu = x / (np.float32(1) + np.exp(-x))
    return silu / x


def gru_step(w, h, token):
    """rnn.cpp:200-258 (Keras GRU, reset_after=True; gate order z, r, h)."""
    u = h.shape[0]
    x = w["emb"][token]                       # ggml_get_rows(embeddings, input_id)          :200
    mx = x @ w["W"] + w["b"][0]               # mul_mat(cell_kernel_t, x) + bias row 0         :203-209
    mh = h @ w["U"] + w["b"][1]               # mul_mat(cell_recurrent_kernel_t, h) + bias 1   :217-225
    z = _sigmoid_like_reference(mx[:u] + mh[:u])              # :231
    r = _sigmoid_like_reference(mx[u:2 * u] + mh[u:2 * u])    # :232
    hh = np.tanh(mx[2 * u:] + r * mh[2 * u:])                 # :235-236
    h2 = (z * h + (np.float32(1) - z) * hh).astype(np.float32)  # :239-250
    logits = h2 @ w["D"] + w["c"]             # :252-258
    return h2, logits.astype(np.float32)


def generate(w, prompt: str, steps: int = 200):
    """inference(), rnn.cpp:266-314: feed the prompt, then feed back the greedy argmax; returns all token ids."""
    char2id = {c: i for i, c in enumerate(VOCAB)}
    ids = [char2id.get(c, char2id["\t"]) for c in prompt]
    h = np.zeros(w["U"].shape[0], np.float32)   # the reference leaves the state uninitialised (:283); zero here (App. C #12)
    nxt = None
    margins = []
    for i in range(steps):
        if i < len(ids):
            tok = ids[i]
        else:
            ids.append(nxt)
            tok = nxt
        h, logits = gru_step(w, h, tok)
        nxt = int(np.argmax(logits))
        srt = np.sort(logits)
        margins.append(float(srt[-1] - srt[-2]))
    return ids, margins


def generate_batch(w, first_tokens, steps):
    """B independent streams (BASELINE.json config 5): stream b starts from first_tokens[b] with a zero state and feeds
    back its own greedy argmax.  Returns tokens [steps, B], margins [steps, B] (top-1 minus top-2 logit) and final state."""
    ids = np.asarray(first_tokens, dtype=np.int64)
    B, u = ids.shape[0], w["U"].shape[0]
    h = np.zeros((B, u), np.float32)
    toks = np.empty((steps, B), np.int32)
    margins = np.empty((steps, B), np.float32)
    for t in range(steps):
        x = w["emb"][ids]
        mx = x @ w["W"] + w["b"][0]
        mh = h @ w["U"] + w["b"][1]
        z = _sigmoid_like_reference(mx[:, :u] + mh[:, :u])
        r = _sigmoid_like_reference(mx[:, u:2 * u] + mh[:, u:2 * u])
        hh = np.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
        h = (z * h + (np.float32(1) - z) * hh).astype(np.float32)
        logits = (h @ w["D"] + w["c"]).astype(np.float32)
        ids = logits.argmax(1)
        srt = np.sort(logits, axis=1)
        toks[t] = ids
        margins[t] = srt[:, -1] - srt[:, -2]
    return toks, margins, h
