"""BASELINE.json configs 2/3: BERT-base (transformers BertModel(BertConfig()) defaults: 12x768,
vocab 30522; 109,482,240 parameters, 199 tensors), random init, seq 512, synthetic token ids."""
import torch


def bert_base():
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    return BertModel(BertConfig())


def batch(rank: int, n: int = 16, seq: int = 512):
    """token ids ~ U{0..30521}[n, seq] from Generator(seed = 4321 + rank)."""
    gen = torch.Generator().manual_seed(4321 + rank)
    return torch.randint(0, 30522, (n, seq), generator=gen)


def loss_fn(model, ids):
    """A scalar that reaches every parameter (encoder, embeddings and pooler): synthetic objective
    for throughput runs — mean square of the final hidden states plus of the pooled output."""
    out = model(input_ids=ids)
    return out.last_hidden_state.float().pow(2).mean() + out.pooler_output.float().pow(2).mean()
