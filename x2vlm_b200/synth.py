"""Seeded synthetic pre-training batches with the layouts the reference's data pipeline produces
(dataset/pretrain_dataset.py:36-130 masking, :595-660 region masks / collate).

Everything is generated on CPU with a torch.Generator so the same seed gives the same batch on any box.
"""
import math

import torch

CLS, SEP, MASK = 101, 102, 103


def image_text_batch(batch, text_len=40, image_res=224, n_mask=12, seed=1234):
    """Image sub-batch: images, all-valid captions, MLM masking of `n_mask` positions per row."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, image_res, image_res, generator=g)
    text_ids = torch.randint(1000, 30000, (batch, text_len), generator=g)
    text_ids[:, 0] = CLS
    text_ids[:, -1] = SEP
    text_atts = torch.ones(batch, text_len, dtype=torch.long)
    n_mask = min(n_mask, text_len - 2)
    masked_pos = torch.stack([torch.randperm(text_len - 2, generator=g)[:n_mask].sort().values + 1 for _ in range(batch)])
    masked_ids = torch.gather(text_ids, 1, masked_pos)
    text_ids_masked = text_ids.clone()
    text_ids_masked.scatter_(1, masked_pos, MASK)
    return dict(image=image, text_ids=text_ids, text_atts=text_atts, text_ids_masked=text_ids_masked,
                masked_pos=masked_pos, masked_ids=masked_ids)


def region_batch(n_img, n_regions, text_len=40, image_res=224, patch=16, n_mask=12, seed=4321):
    """Region sub-batch (bbox loss): n_img images, n_regions region-text samples grouped by image.
    image_atts follows get_image_attns (dataset/pretrain_dataset.py:595-610): token 0 always 1, patch
    (i,j) is 1 iff it intersects the box; the first sample of every image is the whole image
    (is_image = 1, box [0.5,0.5,1,1], :524-526)."""
    g = torch.Generator().manual_seed(seed)
    b = image_text_batch(n_regions, text_len, image_res, n_mask, seed + 1)
    b["image"] = torch.randn(n_img, 3, image_res, image_res, generator=g)
    idx = torch.randint(0, n_img, (n_regions,), generator=g)
    idx[:n_img] = torch.arange(n_img)  # every image is covered
    idx = idx.sort().values
    np_side = image_res // patch
    atts = torch.zeros(n_regions, 1 + np_side * np_side, dtype=torch.long)
    atts[:, 0] = 1
    bbox = torch.zeros(n_regions, 4)
    is_image = torch.zeros(n_regions, dtype=torch.long)
    seen = set()
    for r in range(n_regions):
        im = int(idx[r])
        if im not in seen:
            seen.add(im)
            is_image[r] = 1
            bbox[r] = torch.tensor([0.5, 0.5, 1.0, 1.0])
            atts[r, :] = 1
            continue
        w, h = (torch.rand(2, generator=g) * 0.6 + 0.2).tolist()
        cx = float(torch.rand(1, generator=g)) * (1 - w) + w / 2
        cy = float(torch.rand(1, generator=g)) * (1 - h) + h / 2
        bbox[r] = torch.tensor([cx, cy, w, h])
        # pixel box (x, y, w, h) -> patch ranges, floor / ceil with at least one patch
        px, py, pw, ph = (cx - w / 2) * image_res, (cy - h / 2) * image_res, w * image_res, h * image_res
        x_min = min(math.floor(px / patch), np_side - 1)
        x_max = max(x_min + 1, min(math.ceil((px + pw) / patch), np_side))
        y_min = min(math.floor(py / patch), np_side - 1)
        y_max = max(y_min + 1, min(math.ceil((py + ph) / patch), np_side))
        pm = torch.zeros(np_side, np_side, dtype=torch.long)
        pm[y_min:y_max, x_min:x_max] = 1  # token index = num_patch * row + col + 1
        atts[r, 1:] = pm.reshape(-1)
    b.update(idx_to_group_img=idx, image_atts=atts, target_bbox=bbox, is_image=is_image)
    return b


def hard_negative_indices(batch, seed=7):
    """Deterministic stand-in for the multinomial draw (never the sample itself)."""
    g = torch.Generator().manual_seed(seed)
    shift_i = torch.randint(1, batch, (batch,), generator=g) if batch > 1 else torch.zeros(batch, dtype=torch.long)
    shift_t = torch.randint(1, batch, (batch,), generator=g) if batch > 1 else torch.zeros(batch, dtype=torch.long)
    ar = torch.arange(batch)
    return (ar + shift_i) % batch, (ar + shift_t) % batch
