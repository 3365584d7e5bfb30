"""Does work that waits for an event recorded in the MIDDLE of a stream of pinned host-to-device
copies start when that event's copies are done, or only when the copies enqueued after it are done?"""
import time, torch
torch.cuda.set_device(0)
n, sz = 32, 3110400
host = [torch.empty(sz, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(sz, dtype=torch.uint8, device="cuda") for _ in range(n)]
up, st = torch.cuda.Stream(), torch.cuda.Stream()
x = torch.zeros(1024, device="cuda")
for trial in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs = []
    with torch.cuda.stream(up):
        for i in range(n):
            dev[i].copy_(host[i], non_blocking=True)
            if i == n // 2 - 1:
                e = torch.cuda.Event(); e.record(up); evs.append(e)
    t_enq = time.perf_counter() - t0
    with torch.cuda.stream(st):
        st.wait_event(evs[0])
        x.add_(1)
    st.synchronize()
    t_half = time.perf_counter() - t0
    up.synchronize()
    t_all = time.perf_counter() - t0
    print("enqueue %.3f ms, work after first half of the copies done at %.3f ms, all copies done at %.3f ms" % (t_enq * 1e3, t_half * 1e3, t_all * 1e3))
