"""Interactive inference demo (reference C26: ``bin/pluto.jl:52-382`` — a notebook with a JOB PICKER table that selects the
results of a finished training run and a WEBCAM widget whose RGBA frame is preprocessed, classified and shown with
its top-3 labels).  The same three pieces as a tiny web app (FastAPI + one static page, no JuliaHub):

* ``GET  /jobs``      the job picker: every checkpoint under ``--weights-root`` (the ``weights/<worker>/<model>_cycle_<n>_<time>.pt``
                      files ``start(..., saveweights=True)`` / ``train(..., checkpoint_path=...)`` write), newest first,
                      with step, time and size — the notebook's ``html_results_table`` (``pluto.jl:52-102``);
* ``POST /select``    load the chosen checkpoint into the model (``pluto.jl:104-130``);
* ``POST /classify``  body = raw RGBA bytes of a ``width x height`` canvas frame (exactly what the notebook's camera widget
                      hands over, ``pluto.jl:338-369``) or an encoded JPEG/PNG; returns the top-k labels
                      (``preprocess -> forward -> softmax -> maxk``, ``pluto.jl:371-382``);
* ``GET  /``          the page: job table with radio buttons, a ``getUserMedia`` webcam preview, a capture button and the
                      prediction list.

    python -m fluxdist.bin.demo_server --weights-root weights --model resnet34 --port 8000
"""
from __future__ import annotations

import argparse
import glob
import io
import os
import sys
import time
from typing import Any, Dict, List, Optional

PAGE = """<!doctype html><html><head><meta charset="utf-8"><title>fluxdist demo</title>
<style>body{font-family:sans-serif;margin:2em}td,th{padding:.2em .8em;text-align:left}video,canvas{border:1px solid #888}</style>
</head><body><h2>fluxdist inference demo</h2>
<h3>1. pick a training run</h3><table id="jobs"><thead><tr><th></th><th>time</th><th>step</th><th>file</th></tr></thead><tbody></tbody></table>
<p id="sel">no checkpoint selected</p>
<h3>2. capture</h3><video id="v" width="320" height="240" autoplay muted></video>
<canvas id="c" width="320" height="240" style="display:none"></canvas><br>
<button id="snap">classify frame</button> &nbsp; or upload: <input type="file" id="up" accept="image/*">
<h3>3. predictions</h3><ol id="preds"></ol>
<script>
const tb=document.querySelector('#jobs tbody');
fetch('jobs').then(r=>r.json()).then(js=>{for(const j of js){const tr=document.createElement('tr');
 tr.innerHTML=`<td><input type=radio name=run data-path="${j.path}"></td><td>${j.time}</td><td>${j.step}</td><td>${j.name}</td>`;tb.appendChild(tr)}
 for(const r of tb.querySelectorAll('input'))r.oninput=()=>fetch('select',{method:'POST',headers:{'content-type':'application/json'},
   body:JSON.stringify({path:r.dataset.path})}).then(x=>x.json()).then(x=>document.getElementById('sel').textContent='loaded '+x.loaded)});
const v=document.getElementById('v'),c=document.getElementById('c');
if(navigator.mediaDevices)navigator.mediaDevices.getUserMedia({video:true}).then(s=>v.srcObject=s).catch(()=>{});
function show(p){const ol=document.getElementById('preds');ol.innerHTML='';for(const x of p.topk){const li=document.createElement('li');
 li.textContent=`${x.label}  (${(100*x.p).toFixed(1)} %)`;ol.appendChild(li)}}
document.getElementById('snap').onclick=()=>{const g=c.getContext('2d');g.drawImage(v,0,0,c.width,c.height);
 const d=g.getImageData(0,0,c.width,c.height);fetch(`classify?width=${c.width}&height=${c.height}`,{method:'POST',body:d.data}).then(r=>r.json()).then(show)};
document.getElementById('up').onchange=e=>{fetch('classify',{method:'POST',body:e.target.files[0]}).then(r=>r.json()).then(show)};
</script></body></html>"""


class DemoState:
    """Model + label table + the currently selected run."""

    def __init__(self, model_name: str = "resnet34", nclasses: int = 1000, weights_root: str = "weights",
                 labels: Optional[str] = None, device: Optional[str] = None):
        import torch

        from .driver import build_model

        self.weights_root = weights_root
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model = build_model(model_name, nclasses).eval().to(self.device)
        self.names: Dict[int, str] = {}
        if labels and os.path.exists(labels):
            with open(labels) as f:
                for i, line in enumerate(f):
                    self.names[i] = line.strip().partition(" ")[2]
        self.loaded: Optional[str] = None

    # ---- job picker (pluto.jl:52-102)
    def jobs(self) -> List[Dict[str, Any]]:
        import torch

        out = []
        for p in sorted(glob.glob(os.path.join(self.weights_root, "**", "*.pt"), recursive=True), key=os.path.getmtime,
                        reverse=True):
            step = None
            try:
                step = int(torch.load(p, map_location="cpu", weights_only=False).get("step", 0))
            except Exception:
                pass
            out.append({"path": p, "name": os.path.relpath(p, self.weights_root), "step": step,
                        "time": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(os.path.getmtime(p))),
                        "bytes": os.path.getsize(p)})
        return out

    def select(self, path: str) -> str:
        from ..utils import load_checkpoint

        root = os.path.realpath(self.weights_root)
        real = os.path.realpath(path)
        if not real.startswith(root + os.sep):
            raise ValueError("checkpoint outside the weights root")
        load_checkpoint(real, self.model)
        self.loaded = os.path.relpath(real, root)
        return self.loaded

    # ---- capture -> preprocess -> forward -> top-k (pluto.jl:338-382)
    def classify(self, payload: bytes, width: Optional[int] = None, height: Optional[int] = None, k: int = 3):
        import numpy as np

        from .. import data as D
        from ..utils import maxk
        from .infer_demo import rgba_bytes_to_rgb

        if width and height and len(payload) == width * height * 4:
            img = rgba_bytes_to_rgb(payload, width, height)
        else:
            from PIL import Image

            img = np.asarray(Image.open(io.BytesIO(payload)).convert("RGB"))
        probs = D.forward(self.model, img)
        return [{"index": int(i), "label": self.names.get(int(i), f"class {int(i)}"), "p": float(p)} for i, p in maxk(probs, k)]


def create_app(state: DemoState):
    from fastapi import FastAPI, Request
    from fastapi.responses import HTMLResponse, JSONResponse

    app = FastAPI(title="fluxdist demo")

    @app.get("/", response_class=HTMLResponse)
    def page():
        return PAGE

    @app.get("/jobs")
    def jobs():
        return state.jobs()

    @app.post("/select")
    async def select(req: Request):
        body = await req.json()
        try:
            return {"loaded": state.select(body["path"])}
        except Exception as e:  # noqa: BLE001
            return JSONResponse({"error": str(e)}, status_code=400)

    @app.post("/classify")
    async def classify(req: Request, width: Optional[int] = None, height: Optional[int] = None, k: int = 3):
        return {"run": state.loaded, "topk": state.classify(await req.body(), width, height, k)}

    return app


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--weights-root", default="weights")
    ap.add_argument("--model", default="resnet34")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--labels", default=None, help="LOC_synset_mapping.txt")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    args = ap.parse_args(argv)
    import uvicorn

    state = DemoState(args.model, args.classes, args.weights_root, args.labels)
    uvicorn.run(create_app(state), host=args.host, port=args.port)
    return 0


if __name__ == "__main__":
    sys.exit(main())
