"""Host -> device input pipeline (replaces the per-step feed_dict stall of gan_resnet.py:919-947 /
mnist/model.py:337-372): the next step's inputs are staged in pinned host memory and copied to the device on a dedicated copy
stream WHILE the current step runs; at the step boundary the compute stream only waits for the copy's event and moves the
staged bytes into the program's (graph-captured, fixed-address) input buffers with device-to-device copies.  Two staging
slots alternate, so a slot is never overwritten before the step that consumes it has been enqueued."""
import numpy as np
import torch


class Prefetcher:
    def __init__(self, program, device, slots=2):
        """program: graph.Program whose .inputs {name: Tensor} are fed."""
        self.prog, self.device = program, torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [self._make_slot() for _ in range(slots)]
        self.head = 0           # next slot to fill
        self.ready = []         # filled slots in order

    def _make_slot(self):
        host = {n: torch.empty(t.data.shape, dtype=t.data.dtype).pin_memory() for n, t in self.prog.inputs.items()}
        dev = {n: torch.empty_like(t.data) for n, t in self.prog.inputs.items()}
        return {'host': host, 'dev': dev, 'event': torch.cuda.Event(), 'consumed': torch.cuda.Event(), 'names': [], 'used': False}

    def prefetch(self, **feeds):
        """Convert + stage `feeds` and start their host->device copy; returns immediately."""
        assert len(self.ready) < len(self.slots), 'all staging slots are in flight: commit() first'
        slot = self.slots[self.head]
        self.head = (self.head + 1) % len(self.slots)
        if slot['used']:
            slot['consumed'].synchronize()          # the step that read this slot's device buffers has been enqueued AND run
        slot['names'] = []
        for name, src in feeds.items():
            if src is None:
                continue
            h = slot['host'][name]
            src = torch.as_tensor(np.asarray(src) if not torch.is_tensor(src) else src)
            h.copy_(src.reshape(-1).to(h.dtype))    # host-side conversion (e.g. int64 pixels -> uint8) into pinned memory
            slot['names'].append(name)
        with torch.cuda.stream(self.copy_stream):
            for name in slot['names']:
                slot['dev'][name].copy_(slot['host'][name], non_blocking=True)
            slot['event'].record(self.copy_stream)
        slot['used'] = True
        self.ready.append(slot)

    def commit(self):
        """On the current (compute) stream: wait for the oldest staged set and move it into the program's input buffers."""
        slot = self.ready.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot['event'])
        for name in slot['names']:
            self.prog.inputs[name].data.copy_(slot['dev'][name], non_blocking=True)
        slot['consumed'].record(cur)
