import sys, os
sys.path.insert(0, os.getcwd())
import torch, ctypes as C
import smelter_b200 as s
from smelter_b200 import _ffi as F
import bench
dev = torch.device("cuda", 0)
bg = s.RGBAColor(0x33,0x33,0x33,255)
def streams(n): return [s.InputStreamComponent(input_id=f"input_{i}") for i in range(1,n+1)]
def scene(radius, overlay, wrap_view=True):
    kids = [s.RescalerComponent(child=c, border_radius=s.BorderRadius.new_with_radius(radius)) for c in streams(16)]
    tiles = s.TilesComponent(children=kids, background_color=bg)
    ch = [tiles]
    if overlay:
        ch.append(s.ViewComponent(position=s.Position.Absolute(width=1600.0, height=360.0, left=1120.0, bottom=120.0),
                  background_color=s.RGBAColor(16,32,160,112), border_radius=s.BorderRadius.new_with_radius(48.0)))
    return s.ViewComponent(background_color=bg, children=ch) if wrap_view else tiles
W,H,n,iw,ih = 3840,2160,16,3840,2160
frames = [bench.synth_planes_torch(torch, dev, iw, ih, 100+i) for i in range(n)]
ids = [f"input_{i}".encode() for i in range(1,n+1)]
out_y = torch.empty((H,W),dtype=torch.uint8,device=dev); out_uv = torch.empty((H//2,W//2,2),dtype=torch.uint8,device=dev)
for name,sc in [("full",scene(32.0,True)),("no_overlay",scene(32.0,False)),("no_radius",scene(0.0,True)),("plain",scene(0.0,False)),("tiles_only",scene(0.0,False,False))]:
    r = s.Renderer(s.RendererOptions())
    for b in ids: r.register_input(b.decode())
    r.update_scene("output_1", s.Resolution(W,H), s.OutputFrameFormat.Nv12WgpuTexture, sc)
    arr = (F.InputFrame*n)()
    for i in range(n):
        arr[i].input_id = ids[i]; arr[i].format = F.FRAME_NV12; arr[i].width, arr[i].height = iw, ih; arr[i].mem_kind = F.MEM_DEVICE
        arr[i].planes[0], arr[i].planes[1] = frames[i][0].data_ptr(), frames[i][1].data_ptr()
    o = (F.OutputFrame*1)(); o[0].output_id=b"output_1"; o[0].mem_kind=F.MEM_DEVICE; o[0].planes[0], o[0].planes[1] = out_y.data_ptr(), out_uv.data_ptr()
    for k in range(3): r.render_raw(0, arr, n, o, 1)
    r.set_profiling(True)
    for k in range(10): r.render_raw(0, arr, n, o, 1)
    kt = r.kernel_times()
    print(name, {k:round(v[0]/v[1],4) for k,v in kt.items() if v[1]})
