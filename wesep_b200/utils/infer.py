"""Whole-utterance inference + scoring loop — reference wesep/bin/infer.py:108-181.

The reference walks a DataLoader of 2-speaker mixtures (batch_size 1 = two rows sharing one
mixture, tse_collate_fn_2spk), runs the model under no_grad, peak-normalises, moves everything to
numpy and scores each row with cal_SISNRi.  Here the batch dict keeps its wire format
(wav_mix / wav_targets / spk_embeds / spk / key) but may hold any number of rows; the forward,
the peak rule and SI-SNR / SI-SNRi all stay on the device and ONE device->host read per batch
returns the per-row scores (and the waves only if a sink asks for them).
"""
import torch

from wesep_b200.utils.score import score_batch


@torch.no_grad()
def infer_batch(model, batch, device=None):
    """One iteration of infer.py:109-172.  Returns (waves [n, T] device tensor after the peak rule,
    sisnr [n], sisnri [n]) — device tensors, no host sync."""
    if device is None:
        device = next(model.parameters()).device
    features = batch["wav_mix"].to(device, non_blocking=True).float()
    targets = batch["wav_targets"].to(device, non_blocking=True).float()
    enroll = batch["spk_embeds"].to(device, non_blocking=True).float()
    outputs = model(features, enroll)
    if isinstance(outputs, (list, tuple)):
        outputs = outputs[0]                                            # infer.py:121-122
    waves, sisnr, sisnri, _ = score_batch(outputs, targets, features, peak_norm=True)
    return waves, sisnr, sisnri


def run_inference(model, batches, device=None, sink=None, log=None):
    """The loop of infer.py:108-181 over an iterable of batch dicts.  `sink(name, wave_numpy)` is
    called per row when given (infer.py:131-142 writes `Utt{cnt}-{key}-T{spk}.wav` there);
    `log(str)` receives the per-utterance line of infer.py:154-156.  Returns the summary the
    reference logs at the end: dict(count, sisnr, sisnri, accept) with accept = #rows whose
    SI-SNRi exceeds 1 dB (infer.py:160-161)."""
    was_training = model.training
    model.eval()
    total_sisnr = total_sisnri = 0.0
    cnt = accept = 0
    rows = []
    try:
        for batch in batches:
            waves, sisnr, sisnri = infer_batch(model, batch, device)
            host = torch.stack([sisnr, sisnri]).cpu()                   # the one sync of this batch
            keys = batch.get("key", None)
            spks = batch.get("spk", None)
            wav_host = waves.cpu().numpy() if sink is not None else None
            for r in range(host.shape[1]):
                s, d = float(host[0, r]), float(host[1, r])
                key = keys[r] if keys is not None else str(cnt)
                spk = spks[r] if spks is not None else ""
                if log is not None:
                    log("Num={} | Utt={} | Target speaker={} | SI-SNR={:.2f} | SI-SNRi={:.2f}".format(cnt + 1, key, spk, s, d))
                if sink is not None:
                    sink(f"Utt{cnt + 1}-{key}-T{spk}.wav", wav_host[r])
                rows.append((key, spk, s, d))
                total_sisnr += s
                total_sisnri += d
                cnt += 1
                if d > 1:
                    accept += 1
    finally:
        model.train(was_training)
    return dict(count=cnt, sisnr=total_sisnr / max(cnt, 1), sisnri=total_sisnri / max(cnt, 1), accept=accept, rows=rows)
