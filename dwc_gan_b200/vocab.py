"""CelebA vocabulary with the reference's Vocab API (vocab.py:168-235): 4 specials + the word list = ids used by
the text encoder's embedding table (and by pretrained-embedding lookup)."""
import numpy as np

PAD, BOS, EOS, UNK = '<_>', '<bos>', '<eos>', '<unk>'

_CELEBA_WORDS = "".join([
    'black|blond|brown|male|female|gender|smile|smiling|happy|unsmile|unsmiling|young|younger|old|',
    'older|age|big|glasses|eyeglasses|sunglasses|beard|beards|make|change|translate|modify|reverse|',
    'inverse|increase|add|decrease|reduce|boy|man|gentleman|sir|woman|lady|miss|girl|moustache|',
    'whiskers|delighted|laugh|unhappy|serious|smileless|solemn|less|more|attractive|attractiveness|',
    'do|not|nothing|anything|everything|keep|unchanged|his|him|it|the|its|her|face|wear|put|on|with|',
    'remove|take|off|without|no|to|into|and|unknown|,|.|color|colour|hair|from|be|a|an|this|wearing|',
    'gray|left|right|but|blonde| |?|!',
]).split("|")


class Vocab(object):
    def __init__(self, dataset='CelebA', with_SE=True):
        if dataset != 'CelebA':
            raise NotImplementedError("only the CelebA vocabulary is on the hot path")
        specials = [PAD, BOS, EOS, UNK] if with_SE else [PAD, UNK]
        self.itos = specials + list(_CELEBA_WORDS)
        self.stoi = {w: i for i, w in enumerate(self.itos)}
        self._size = len(self.stoi)
        self._padding_idx = self.stoi[PAD]
        self._unk_idx = self.stoi[UNK]
        self._start_idx = self.stoi.get(BOS, -1)
        self._end_idx = self.stoi.get(EOS, -1)

    def idx2token(self, x):
        return [self.idx2token(i) for i in x] if isinstance(x, list) else self.itos[x]

    def token2idx(self, x):
        return [self.token2idx(i) for i in x] if isinstance(x, list) else self.stoi.get(x, self._unk_idx)

    def random_sample(self):
        return self.idx2token(1 + np.random.randint(self._size - 1))

    size = property(lambda self: self._size)
    padding_idx = property(lambda self: self._padding_idx)
    unk_idx = property(lambda self: self._unk_idx)
    start_idx = property(lambda self: self._start_idx)
    end_idx = property(lambda self: self._end_idx)


def ListsToTensor(xs, vocab, with_S=True, with_E=True, mx_len=50):
    """Token lists -> ([B, mx_len] id array padded with PAD, lengths) (vocab.py:220-235)."""
    xs = [list(x[:mx_len]) for x in xs]
    extra = int(with_S) + int(with_E)
    lens = [len(x) + extra for x in xs]
    rows = []
    for x, n in zip(xs, lens):
        row = ([vocab.start_idx] if with_S else []) + [vocab.token2idx(w) for w in x] + \
              ([vocab.end_idx] if with_E else [])
        rows.append(row + [vocab.padding_idx] * (mx_len - n))
    return np.array(rows), np.array([max(1, n) for n in lens])
