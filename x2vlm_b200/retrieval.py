"""Image-text retrieval scoring on the B200-native encoders: ITC all-pairs similarities, then ITM re-ranking of the
top-k candidates of every image (i2t) and every caption (t2i) through the fusion layers.

Mirrors the reference's Retrieval.py:evaluation (:71-157) — same inputs (a model exposing get_vision_embeds /
get_text_embeds / get_features / get_cross_embeds / itm_head), same outputs (two score matrices filled with -100
outside the re-ranked candidates), same rank partition + SUM all-reduce across processes — with the two
sequential loops (one fusion call of k_test sequences per image, then per caption; 6 000 calls for the 1k/5k test
set) replaced by batched calls:

  * i2t: `rows_per_call` images per fusion call; the k_test captions of one image all point at that image's K/V
    through `encoder_kv_index`, so the image K/V projection runs once per image per layer instead of k_test times
    (the reference repeats the image tensor k_test times, Retrieval.py:127);
  * t2i: `rows_per_call` captions per call; the candidate images of the whole call are de-duplicated and each
    (caption, image) sequence indexes its image's K/V.
"""
import torch
import torch.distributed as dist


def _rank_slice(n):
    """Rows [start, end) of this process (Retrieval.py:120-123: step = n // world + 1)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    step = n // world + 1
    start = rank * step
    return start, min(n, start + step), world


@torch.no_grad()
def encode(model, images, text_ids, text_atts, image_bs=64, text_bs=256):
    """Features of the whole test set (Retrieval.py:81-116): per-token embeddings and the normalised ITC features."""
    text_embeds, text_feats = [], []
    for i in range(0, text_ids.shape[0], text_bs):
        e = model.get_text_embeds(text_ids[i:i + text_bs], text_atts[i:i + text_bs])
        text_embeds.append(e)
        text_feats.append(model.get_features(text_embeds=e))
    image_embeds, image_feats = [], []
    for i in range(0, images.shape[0], image_bs):
        e, _ = model.get_vision_embeds(images[i:i + image_bs])
        image_embeds.append(e)
        image_feats.append(model.get_features(image_embeds=e))
    return torch.cat(image_embeds), torch.cat(image_feats), torch.cat(text_embeds), torch.cat(text_feats)


@torch.no_grad()
def rerank(model, image_embeds, text_embeds, text_atts, sims_matrix, k_test, rows_per_call=8, reduce=True):
    """ITM scores of the top-k candidates (Retrieval.py:118-151).  Returns (score_matrix_i2t [n_img, n_txt],
    score_matrix_t2i [n_txt, n_img]) on the device, -100 where a pair was not re-ranked."""
    dev = sims_matrix.device
    n_img, n_txt = sims_matrix.shape
    k_i2t, k_t2i = min(k_test, n_txt), min(k_test, n_img)
    score_i2t = torch.full((n_img, n_txt), -100.0, device=dev)
    score_t2i = torch.full((n_txt, n_img), -100.0, device=dev)

    def itm(cross):
        return model.itm_head(cross[:, 0, :].float())[:, 1]

    start, end, world = _rank_slice(n_img)
    for r0 in range(start, end, rows_per_call):
        r1 = min(end, r0 + rows_per_call)
        topk_idx = sims_matrix[r0:r1].topk(k=k_i2t, dim=1).indices                    # [rows, k] caption ids
        flat = topk_idx.reshape(-1)
        enc = image_embeds[r0:r1]
        kv_index = torch.arange(r1 - r0, device=dev, dtype=torch.int32).repeat_interleave(k_i2t)
        enc_atts = torch.ones(flat.numel(), enc.shape[1], dtype=torch.long, device=dev)
        out = model.get_cross_embeds(enc, enc_atts, text_embeds=text_embeds[flat], text_atts=text_atts[flat],
                                     encoder_kv_index=kv_index)
        score_i2t[r0:r1].scatter_(1, topk_idx, itm(out).view(r1 - r0, k_i2t))

    start, end, _ = _rank_slice(n_txt)
    sims_t = sims_matrix.t()
    for r0 in range(start, end, rows_per_call):
        r1 = min(end, r0 + rows_per_call)
        topk_idx = sims_t[r0:r1].topk(k=k_t2i, dim=1).indices                         # [rows, k] image ids
        uniq, inv = torch.unique(topk_idx.reshape(-1), return_inverse=True)             # K/V once per distinct image
        rows = torch.arange(r0, r1, device=dev).repeat_interleave(k_t2i)
        enc = image_embeds[uniq]
        enc_atts = torch.ones(rows.numel(), enc.shape[1], dtype=torch.long, device=dev)
        out = model.get_cross_embeds(enc, enc_atts, text_embeds=text_embeds[rows], text_atts=text_atts[rows],
                                     encoder_kv_index=inv.to(torch.int32))
        score_t2i[r0:r1].scatter_(1, topk_idx, itm(out).view(r1 - r0, k_t2i))

    if reduce and world > 1:
        combine_rank_rows(score_i2t)
        combine_rank_rows(score_t2i)
    return score_i2t, score_t2i


def combine_rank_rows(scores):
    """Every process scored the rows of its _rank_slice into a matrix pre-filled with -100; the reference SUM-all-reduces
    those matrices as they are (Retrieval.py:145-148), so with W processes every entry carries an extra -100·(W-1)
    (a constant shift per matrix: rankings are unchanged).  Same call here — the score matrices are the reference's,
    entry for entry, at any world size."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(scores, op=dist.ReduceOp.SUM)
    return scores


@torch.no_grad()
def evaluation(model, images, text_ids, text_atts, k_test=128, image_bs=64, text_bs=256, rows_per_call=8):
    """Retrieval.py:evaluation on tensors already tokenised / pre-processed: returns (score_matrix_i2t,
    score_matrix_t2i, sims_matrix)."""
    model.eval()
    image_embeds, image_feats, text_embeds, text_feats = encode(model, images, text_ids, text_atts, image_bs, text_bs)
    sims_matrix = image_feats @ text_feats.t()
    s_i2t, s_t2i = rerank(model, image_embeds, text_embeds, text_atts, sims_matrix, k_test, rows_per_call)
    return s_i2t, s_t2i, sims_matrix


def itm_eval(scores_i2t, scores_t2i, txt2img, img2txt):
    """Recall@{1,5,10} both ways (Retrieval.py:160-209): rank of the best-scoring ground-truth caption / image."""
    import numpy as np
    s_i2t = scores_i2t.detach().float().cpu().numpy() if isinstance(scores_i2t, torch.Tensor) else scores_i2t
    s_t2i = scores_t2i.detach().float().cpu().numpy() if isinstance(scores_t2i, torch.Tensor) else scores_t2i
    ranks = np.zeros(s_i2t.shape[0])
    for index, score in enumerate(s_i2t):
        inds = np.argsort(score)[::-1]
        ranks[index] = min(np.where(inds == i)[0][0] for i in img2txt[index])
    tr = [100.0 * len(np.where(ranks < k)[0]) / len(ranks) for k in (1, 5, 10)]
    ranks = np.zeros(s_t2i.shape[0])
    for index, score in enumerate(s_t2i):
        inds = np.argsort(score)[::-1]
        ranks[index] = np.where(inds == txt2img[index])[0][0]
    ir = [100.0 * len(np.where(ranks < k)[0]) / len(ranks) for k in (1, 5, 10)]
    return {'txt_r1': tr[0], 'txt_r5': tr[1], 'txt_r10': tr[2], 'txt_r_mean': sum(tr) / 3, 'img_r1': ir[0], 'img_r5': ir[1],
            'img_r10': ir[2], 'img_r_mean': sum(ir) / 3, 'r_mean': (sum(tr) + sum(ir)) / 6}
