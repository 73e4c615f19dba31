"""TEST INFRASTRUCTURE — CPU restatement (numpy + torch CPU) of the reference's video-level driver, SURVEY §8(f)
rank 4: the sliding-window clip scheduler, mirror padding, hole compositing and 0.5/0.5 blending of
``/root/reference/test.py:37-52,132-179``.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU leg may
import this module; the product path (``e2fgvi_b200/video.py``) never does.

Pinned against the UNMODIFIED ``test.py`` run in the build container (``oracle/gen_golden_video.py`` executes its
``main_worker`` with matplotlib stubbed and ``cv2.VideoWriter`` captured): ``tests/golden/video_*.npz`` holds the
composited frames that script wrote, and ``tests/test_oracle.py`` checks this restatement reproduces them bit for bit.
"""
import numpy as np
import torch


def get_ref_index(f, neighbor_ids, length, ref_length=10, num_ref=-1):
    """Non-local reference frame ids for the window centred on ``f`` (test.py:37-52).  The reference reads
    ``ref_length`` / ``num_ref`` from module globals (test.py:31-32); note its ``len(ref_index) > num_ref`` test lets
    ``num_ref + 1`` ids through."""
    ref_index = []
    if num_ref == -1:
        for i in range(0, length, ref_length):
            if i not in neighbor_ids:
                ref_index.append(i)
    else:
        start_idx = max(0, f - ref_length * (num_ref // 2))
        end_idx = min(length, f + ref_length * (num_ref // 2))
        for i in range(start_idx, end_idx + 1, ref_length):
            if i not in neighbor_ids:
                if len(ref_index) > num_ref:
                    break
                ref_index.append(i)
    return ref_index


def window_schedule(video_length, neighbor_stride=5, ref_length=10, num_ref=-1):
    """[(f, neighbor_ids, ref_ids)] for f = 0, stride, 2*stride, ... (test.py:146-151)."""
    out = []
    for f in range(0, video_length, neighbor_stride):
        neighbor_ids = [i for i in range(max(0, f - neighbor_stride), min(video_length, f + neighbor_stride + 1))]
        out.append((f, neighbor_ids, get_ref_index(f, neighbor_ids, video_length, ref_length, num_ref)))
    return out


def dilate_cross(mask, iterations=4):
    """``cv2.dilate(m, MORPH_CROSS 3x3, iterations=4)`` of a 0/1 uint8 map (test.py:63-65) — part of the reference's
    mask IO, restated so that synthetic masks can be fed to both sides of a parity test without cv2."""
    m = mask.astype(bool)
    for _ in range(iterations):
        d = m.copy()
        d[1:] |= m[:-1]
        d[:-1] |= m[1:]
        d[:, 1:] |= m[:, :-1]
        d[:, :-1] |= m[:, 1:]
        m = d
    return m.astype(np.uint8)


def inpaint_video(model, frames_u8, masks_u8, neighbor_stride=5, ref_length=10, num_ref=-1, pred_hook=None):
    """test.py:132-179.  ``frames_u8`` (N,H,W,3) uint8 RGB, ``masks_u8`` (N,H,W) uint8 0/1 (already dilated),
    ``model(masked_imgs[1,t,3,H',W'], l_t) -> (pred[t,3,H',W'], flows)`` on CPU.  Returns the list of composited
    frames exactly as the reference leaves them in ``comp_frames`` before the final ``astype(np.uint8)``:
    uint8 arrays for frames seen once, float32 for blended ones.  ``pred_hook(window_index, pred)`` lets a test
    substitute / record the network output."""
    n, h, w, _ = frames_u8.shape
    # to_tensors()(frames) * 2 - 1  /  to_tensors()(masks)   (core/utils.py:138-178, test.py:132,139)
    imgs = torch.from_numpy(frames_u8).permute(0, 3, 1, 2).contiguous().float().div(255).unsqueeze(0) * 2 - 1
    masks = torch.from_numpy(masks_u8 * 255).unsqueeze(1).contiguous().float().div(255).unsqueeze(0)
    binary_masks = [np.expand_dims((masks_u8[i] != 0).astype(np.uint8), 2) for i in range(n)]
    frames = [frames_u8[i] for i in range(n)]
    comp_frames = [None] * n
    for wi, (f, neighbor_ids, ref_ids) in enumerate(window_schedule(n, neighbor_stride, ref_length, num_ref)):
        selected_imgs = imgs[:1, neighbor_ids + ref_ids]
        selected_masks = masks[:1, neighbor_ids + ref_ids]
        with torch.no_grad():
            masked_imgs = selected_imgs * (1 - selected_masks)
            h_pad = (60 - h % 60) % 60
            w_pad = (108 - w % 108) % 108
            masked_imgs = torch.cat([masked_imgs, torch.flip(masked_imgs, [3])], 3)[:, :, :, :h + h_pad, :]
            masked_imgs = torch.cat([masked_imgs, torch.flip(masked_imgs, [4])], 4)[:, :, :, :, :w + w_pad]
            pred_imgs, _ = model(masked_imgs, len(neighbor_ids))
            if pred_hook is not None:
                pred_imgs = pred_hook(wi, pred_imgs)
            pred_imgs = pred_imgs[:, :, :h, :w]
            pred_imgs = (pred_imgs + 1) / 2
            pred_imgs = pred_imgs.cpu().permute(0, 2, 3, 1).numpy() * 255
        for i in range(len(neighbor_ids)):
            idx = neighbor_ids[i]
            img = np.array(pred_imgs[i]).astype(np.uint8) * binary_masks[idx] + frames[idx] * (1 - binary_masks[idx])
            if comp_frames[idx] is None:
                comp_frames[idx] = img
            else:
                comp_frames[idx] = comp_frames[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
    return comp_frames


def finalize(comp_frames):
    """The ``comp_frames[f].astype(np.uint8)`` of test.py:195 for every frame -> (N,H,W,3) uint8."""
    return np.stack([c.astype(np.uint8) for c in comp_frames])


# ------------------------------------------------------------------------------------------------------------------
# The same arithmetic cut at the four kernel boundaries of e2fgvi_b200/csrc/video.cu (torch CPU, fp32), so that each
# kernel can be checked bit for bit and the driver's host logic can run on a CPU-only box (tests only).
def prepare_clip(frames, masks, ids, hp, wp):
    """test.py:132,139,152-165 for one window: (N,H,W,3) u8, (N,H,W) u8, ids -> (t,3,hp,wp) fp32."""
    ids = [int(i) for i in ids]
    h, w = frames.shape[1:3]
    imgs = frames[ids].permute(0, 3, 1, 2).contiguous().float().div(255) * 2 - 1
    m = (masks[ids] != 0).float().unsqueeze(1)
    x = imgs * (1 - m)
    x = torch.cat([x, torch.flip(x, [2])], 2)[:, :, :hp, :]
    x = torch.cat([x, torch.flip(x, [3])], 3)[:, :, :, :wp]
    return x.contiguous()


def compose(pred, frames, masks, ids, n_local):
    """test.py:167-174: (>=n_local,3,hp,wp) fp32 -> (n_local,H,W,3) u8."""
    ids = [int(i) for i in ids][:n_local]
    h, w = frames.shape[1:3]
    p = (pred[:n_local, :, :h, :w] + 1) / 2
    p = (p.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
    b = (masks[ids] != 0).numpy().astype(np.uint8)[..., None]
    return torch.from_numpy(p * b + frames[ids].numpy() * (1 - b))


def blend(img, ids, first, comp):
    """test.py:175-179 on an fp32 canvas (N,H,W,3); uint8 -> fp32 is exact, so keeping first-seen frames as fp32 is
    equivalent to the reference's uint8-then-float32 bookkeeping."""
    for k, idx in enumerate([int(i) for i in ids]):
        v = img[k].float()
        comp[idx] = v if int(first[k]) else comp[idx] * 0.5 + v * 0.5
    return comp


def finalize_canvas(comp):
    return torch.from_numpy(comp.numpy().astype(np.uint8))
