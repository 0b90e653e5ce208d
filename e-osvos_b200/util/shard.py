"""Video/object sharding across the GPUs of one box: the unit of work is one (video, object)
under multi_object='single_id' (reference evaluate.py:111,132 -- a fresh theta_0 per object; objects only meet
at the host-side merge :323-326), so ranks exchange NOTHING on the data path.  Longest-processing-time-first
over a FLOP model of e-OSVOS(-OnA): 3 images x 1.2 TFLOP per fine-tune iteration, 0.42 TFLOP per inference frame.
The reference itself runs one dataset per GPU, sequences serially (train_meta.py:134-146,175-186)."""


def ona_schedule(num_frames, num_epochs_eval, online_adapt_step=0, online_adapt_epochs=10):
    """Rounds of reference evaluate.py:140-317 for one object: list of (round, iters, frame_min, frame_max)
    with frames [frame_min, frame_max) inferred after the round's fine-tuning."""
    T = num_frames
    if online_adapt_step:
        step, n_rounds = online_adapt_step, len(range(1, T, online_adapt_step))
    else:
        step, n_rounds = T, 1
    out, range_max = [], 1
    for k in range(n_rounds):
        range_min = range_max
        range_max = min(range_max + step, T)
        out.append((k, num_epochs_eval if k == 0 else online_adapt_epochs, range_min, range_max))
        if range_max == T:
            break
    return out


def unit_cost(num_frames, num_epochs_eval, online_adapt_step=0, online_adapt_epochs=10, batch=3):
    sched = ona_schedule(num_frames, num_epochs_eval, online_adapt_step, online_adapt_epochs)
    iters = sum(s[1] for s in sched)
    frames = sum(s[3] - s[2] for s in sched)
    return iters * batch * 1.2 + frames * 0.42


def shard_units(units, world_size, **cfg):
    """units: list of (video_id, object_id, num_frames).  Returns world_size lists (LPT assignment)."""
    costed = sorted(((unit_cost(u[2], **cfg), i, u) for i, u in enumerate(units)), reverse=True)
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for c, _, u in costed:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        shards[r].append(u)
        loads[r] += c
    return shards, loads
