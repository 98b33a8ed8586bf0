"""
processing_pipeline -- ProcessingPipelineThread with the reference's constructor, Qt signals, file
selection / ordering / naming rules and error surface (processing_pipeline.py:21-273); the per-layer
work (:74-113) and the sliding window of prior masks (:163, :216-223) run on the device through
vsb_engine.StackEngine.  PNG decode / encode stay on host threads (north_star) and overlap with the
device pipeline: decode pool -> pinned buffer -> H2D | kernels | D2H (three CUDA streams) -> encode pool.

PySide6 is optional: on headless boxes a minimal Signal/QThread shim with the same call surface is used.
"""
from __future__ import annotations

import collections
import concurrent.futures
import os
import queue
import re
import shutil
import threading
import time
import traceback

import cv2
import numpy as np

try:  # the GUI toolkit is only needed for the signal plumbing
    from PySide6.QtCore import QThread, Signal
except Exception:  # pragma: no cover - exercised on headless boxes
    class _BoundSignal:
        def __init__(self):
            self._slots = []

        def connect(self, fn):
            self._slots.append(fn)

        def emit(self, *args):
            for fn in list(self._slots):
                fn(*args)

    class Signal:  # descriptor: one bound signal per instance, like Qt
        def __init__(self, *types):
            self._name = None

        def __set_name__(self, owner, name):
            self._name = "_sig_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            sig = obj.__dict__.get(self._name)
            if sig is None:
                sig = obj.__dict__[self._name] = _BoundSignal()
            return sig

    class QThread:
        def __init__(self, *a, **k):
            pass

        def start(self):
            self.run()

        def isRunning(self):
            return False

        def wait(self, *a):
            return True

import lut_manager
import uvtools_wrapper
import vsb_engine as E
import vsb_params as P
from config import Config, ProcessingMode
from logger import Logger

_TRAILING_INT = re.compile(r"(\d+)\.\w+$")
_IMAGE_EXT = (".png", ".bmp", ".tif", ".tiff")


def _layer_number(filename):
    m = _TRAILING_INT.search(filename)
    return int(m.group(1)) if m else float("inf")


def list_layer_files(folder, start_index=None, stop_index=None):
    """Files of a stack in processing order (processing_pipeline.py:141-153): image extensions only,
    sorted by the trailing integer of the name, filtered to [start_index, stop_index]."""
    names = sorted((f for f in os.listdir(folder) if f.lower().endswith(_IMAGE_EXT)), key=_layer_number)
    keep = []
    for f in names:
        n = _layer_number(f)
        if start_index is not None and n < start_index:
            continue
        if stop_index is not None and n > stop_index:
            continue
        keep.append(f)
    return keep


class ProcessingPipelineThread(QThread):
    status_update = Signal(str)
    progress_update = Signal(int)
    error_signal = Signal(str)
    finished_signal = Signal()

    def __init__(self, app_config: Config, max_workers: int):
        super().__init__()
        self.app_config = app_config
        self._is_running = True
        self.error_occurred = False
        self.max_workers = max(1, int(max_workers))
        self.logger = Logger()
        self.run_timestamp = self.logger.run_timestamp
        self.session_temp_folder = ""

    # ---- UVTools host steps (unchanged behaviour, processing_pipeline.py:40-72) -------------------------
    def _run_uvtools_extraction(self) -> str:
        self.status_update.emit("Starting UVTools slice extraction...")
        self.logger.log("Starting UVTools slice extraction.")
        cfg = self.app_config
        self.session_temp_folder = os.path.join(cfg.uvtools_temp_folder, f"{cfg.output_file_prefix}{self.run_timestamp}")
        folder = uvtools_wrapper.extract_layers(cfg.uvtools_path, cfg.uvtools_input_file, self.session_temp_folder)
        self.status_update.emit("UVTools extraction completed.")
        return folder

    def _run_uvtools_repack(self, processed_images_folder: str):
        cfg = self.app_config
        self.status_update.emit("Generating UVTools operation file...")
        uvtop = uvtools_wrapper.generate_uvtop_file(processed_images_folder, self.session_temp_folder, self.run_timestamp)
        self.status_update.emit("Operation file generated.")
        self.status_update.emit("Repacking slice file with processed layers...")
        final_path = uvtools_wrapper.repack_layers(cfg.uvtools_path, cfg.uvtools_input_file, uvtop,
                                                   cfg.uvtools_output_location, cfg.uvtools_temp_folder,
                                                   cfg.output_file_prefix, self.run_timestamp)
        self.status_update.emit(f"Successfully created: {os.path.basename(final_path)}")

    # ---- the device-backed stack loop -------------------------------------------------------------------------
    def _make_engine(self, H, W):
        cfg = self.app_config
        ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params), W, H)
        return E.StackEngine(W, H, cfg, ops, metric=getattr(cfg, "distance_metric", "chamfer5"))

    def _blend_stack_roi(self, input_path, filenames, output_path):
        """ROI_FADE: the tracker is sequential host state, so layers go through one at a time (decode and encode
        still run on the worker pools around it).  processing_pipeline.py:205-223"""
        import xy_blend_processor
        cfg = self.app_config
        total = len(filenames)
        done = 0
        workers = self.max_workers
        engine = None
        try:
            with concurrent.futures.ThreadPoolExecutor(workers) as decoder, \
                    concurrent.futures.ThreadPoolExecutor(workers) as encoder:
                pending = collections.deque()
                it = iter(filenames)

                def top_up():
                    while len(pending) < 2 * workers:
                        name = next(it, None)
                        if name is None:
                            return
                        pending.append((name, decoder.submit(cv2.imread, os.path.join(input_path, name), cv2.IMREAD_GRAYSCALE)))

                def write(path, image):
                    cv2.imwrite(path, image)
                    print(f"Successfully wrote output file: {path}")

                top_up()
                jobs = []
                index = 0
                while pending:
                    if not self._is_running:
                        self.logger.log("Processing stopped by user.")
                        self.status_update.emit("Processing stopped by user.")
                        break
                    name, fut = pending.popleft()
                    top_up()
                    index += 1
                    self.status_update.emit(f"Analyzing {name} ({index}/{total})")
                    gray = fut.result()
                    if gray is None:
                        print(f"Warning: Could not load image at {os.path.join(input_path, name)}")
                        self.status_update.emit(f"Skipping unloadable image: {name}")
                        total = max(1, total - 1)
                        continue
                    if engine is None:
                        H, W = gray.shape
                        ops = list(cfg.xy_blend_pipeline or [])
                        chain = (lambda d: xy_blend_processor.process_xy_pipeline_dev(d, ops)[0]) if ops else None
                        engine = E.RoiStackEngine(W, H, cfg, chain)
                    out = engine.process(gray, _layer_number(name))
                    jobs.append(encoder.submit(write, os.path.join(output_path, name), out))
                    while jobs and (jobs[0].done() or len(jobs) > 2 * workers):
                        jobs.pop(0).result()
                        done += 1
                        self.status_update.emit(f"Completed processing images ({done}/{total})")
                        self.progress_update.emit(int((done / total) * 100))
                for j in jobs:
                    j.result()
                    done += 1
                    self.status_update.emit(f"Completed processing images ({done}/{total})")
                    self.progress_update.emit(int((done / total) * 100))
        finally:
            if engine is not None:
                E.torch().cuda.synchronize()
                engine.close()

    def _blend_stack(self, input_path, filenames, output_path):
        """The stack loop (processing_pipeline.py:141-223) on the device.  With `config.devices` naming several CUDA
        devices the stack is Z-sharded: contiguous slabs, one engine and one host thread per device, each slab first
        re-ingesting the `receding_layers` layers below it as a mask-only halo (no collective)."""
        if self.app_config.blending_mode == ProcessingMode.ROI_FADE:
            return self._blend_stack_roi(input_path, filenames, output_path)
        cfg = self.app_config
        devices = [int(d) for d in (getattr(cfg, "devices", None) or [])]
        self._progress = {"done": 0, "total": len(filenames), "index": 0}
        self._progress_lock = threading.Lock()
        self.io_seconds = {"decode": 0.0, "device_wait": 0.0, "encode": 0.0, "layers": 0}
        if len(devices) <= 1:
            return self._blend_slab(input_path, filenames, [], output_path, devices[0] if devices else None)
        halo = max(0, int(cfg.receding_layers))
        plan = [p for p in E.zshard_plan(len(filenames), len(devices), halo) if p["z1"] > p["z0"]]
        errors = []

        def slab(p):
            try:
                self._blend_slab(input_path, filenames[p["z0"]:p["z1"]], filenames[p["halo0"]:p["z0"]], output_path,
                                 devices[p["rank"]])
            except Exception as exc:  # surfaced by run() exactly like a failed worker future (:179-185)
                errors.append(exc)
                self._is_running = False

        threads = [threading.Thread(target=slab, args=(p,), daemon=True) for p in plan]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def _blend_slab(self, input_path, filenames, halo_files, output_path, device=None):
        prog, lock = self._progress, self._progress_lock
        torch = E.torch()
        if device is not None:
            torch.cuda.set_device(device)
        workers = self.max_workers
        engine = None
        free_bufs: "queue.Queue" = queue.Queue()
        inflight = collections.deque()
        encode_jobs = set()
        io = self.io_seconds

        def finish_encodes(block_all=False):
            if not encode_jobs:
                return
            if block_all:
                ready = concurrent.futures.wait(encode_jobs)[0]
            else:
                ready = {j for j in encode_jobs if j.done()}
            for job in ready:
                encode_jobs.discard(job)
                job.result()
                with lock:
                    prog["done"] += 1
                    done, total = prog["done"], prog["total"]
                self.status_update.emit(f"Completed processing images ({done}/{total})")
                self.progress_update.emit(int((done / total) * 100))

        def decode(path):
            t0 = time.perf_counter()
            img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
            with lock:
                io["decode"] += time.perf_counter() - t0
            return img

        def encode(buf, path, W):
            try:
                t0 = time.perf_counter()
                cv2.imwrite(path, buf[1][:, :W])
                with lock:
                    io["encode"] += time.perf_counter() - t0
                    io["layers"] += 1
                print(f"Successfully wrote output file: {path}")
            finally:
                free_bufs.put(buf)
            return path

        def retire(encoder):
            ticket, buf, path = inflight.popleft()
            t0 = time.perf_counter()
            engine.wait(ticket)
            with lock:
                io["device_wait"] += time.perf_counter() - t0
            encode_jobs.add(encoder.submit(encode, buf, path, engine.out_W))

        try:
            with concurrent.futures.ThreadPoolExecutor(workers) as decoder, \
                    concurrent.futures.ThreadPoolExecutor(workers) as encoder:
                lookahead = 2 * workers                                   # same back-pressure bound as :169
                pending = collections.deque()
                it = iter([(n, True) for n in halo_files] + [(n, False) for n in filenames])

                def top_up():
                    while len(pending) < lookahead:
                        item = next(it, None)
                        if item is None:
                            return
                        name, is_halo = item
                        pending.append((name, is_halo, decoder.submit(decode, os.path.join(input_path, name))))

                top_up()
                while pending:
                    if not self._is_running:
                        self.logger.log("Processing stopped by user.")
                        self.status_update.emit("Processing stopped by user.")
                        break
                    name, is_halo, fut = pending.popleft()
                    top_up()
                    if not is_halo:
                        with lock:
                            prog["index"] += 1
                            index, total = prog["index"], prog["total"]
                        self.status_update.emit(f"Analyzing {name} ({index}/{total})")
                    gray = fut.result()
                    if gray is None:
                        print(f"Warning: Could not load image at {os.path.join(input_path, name)}")
                        if not is_halo:
                            self.status_update.emit(f"Skipping unloadable image: {name}")
                            with lock:
                                prog["total"] = max(1, prog["total"] - 1)
                        continue
                    if engine is None:
                        H, W = gray.shape
                        engine = self._make_engine(H, W)
                        for _ in range(engine.slots + workers):
                            pin_in = torch.empty((H, engine.pitch), dtype=torch.uint8, pin_memory=True)
                            pin_out = torch.empty((engine.out_H, E.pitch_for(engine.out_W)), dtype=torch.uint8, pin_memory=True)
                            free_bufs.put((pin_in.numpy(), pin_out.numpy(), pin_in, pin_out))
                    if gray.shape != (engine.H, engine.W):
                        raise ValueError(f"{name}: layer size {gray.shape[::-1]} differs from the stack's "
                                         f"{(engine.W, engine.H)}")
                    if is_halo:                                           # masks only: thresholded, pushed, no output
                        while inflight:
                            retire(encoder)
                        engine.push_halo(gray)
                        continue
                    finish_encodes()
                    buf = free_bufs.get()
                    np.copyto(buf[0][:, : engine.W], gray)
                    ticket = engine.submit(buf[2].data_ptr(), engine.pitch, buf[3].data_ptr(), E.pitch_for(engine.out_W))
                    inflight.append((ticket, buf, os.path.join(output_path, name)))
                    while len(inflight) >= engine.slots:
                        retire(encoder)
                while inflight:
                    retire(encoder)
                finish_encodes(block_all=True)
        finally:
            # a failure above may leave copies in flight between pinned host buffers and the device: drain the streams
            # before those buffers (and the engine's staging slots) are released
            if engine is not None:
                try:
                    engine.sync()
                finally:
                    engine.close()
            for job in list(encode_jobs):
                try:
                    job.result()
                except Exception:
                    pass

    def run(self):
        self.logger.log("Run started.")
        self.logger.log_config(self.app_config)
        self.status_update.emit("Processing started...")
        cfg = self.app_config
        try:
            if cfg.input_mode == "uvtools":
                input_path = self._run_uvtools_extraction()
                self.logger.log("UVTools extraction completed.")
                output_path = os.path.join(self.session_temp_folder, "Output")
                os.makedirs(output_path, exist_ok=True)
            else:
                input_path, output_path = cfg.input_folder, cfg.output_folder

            filenames = list_layer_files(input_path, cfg.start_index, cfg.stop_index)
            if not filenames:
                msg = "No images found in the specified folder or index range."
                self.logger.log(f"ERROR: {msg}")
                self.error_signal.emit(msg)
                return
            self.logger.log(f"Found {len(filenames)} images to process.")
            if cfg.debug_save:
                note = ("debug_save: the device stack engine keeps intermediates on the GPU and writes no debug_03a/03b "
                        "images; call processing_core.process_z_blending(..., debug_info=...) per layer for the dumps.")
                self.logger.log(note)
                self.status_update.emit(note)
            try:
                self._blend_stack(input_path, filenames, output_path)
            except Exception as exc:
                self.error_occurred = True
                detail = f"An image processing task failed: {exc}\n{traceback.format_exc()}"
                self.logger.log(f"ERROR during image processing task: {detail}")
                self.error_signal.emit(detail)
                self.stop_processing()

            self.logger.log("Stack blending complete.")
            if cfg.input_mode == "uvtools" and not self.error_occurred:
                if self._is_running:
                    self.status_update.emit("All image processing tasks completed.")
                else:
                    self.status_update.emit("Processing stopped by user, repacking completed layers...")
                self.logger.log("Starting UVTools repack.")
                self._run_uvtools_repack(output_path)
                self.logger.log("UVTools repack completed.")
        except Exception as exc:
            self.error_occurred = True
            info = f"A critical error occurred in the processing pipeline: {exc}\n\n{traceback.format_exc()}"
            self.logger.log(f"CRITICAL ERROR in pipeline: {info}")
            self.error_signal.emit(info)
        finally:
            self.logger.log("Run finalizing.")
            if (cfg.input_mode == "uvtools" and cfg.uvtools_delete_temp_on_completion and not self.error_occurred
                    and self.session_temp_folder and os.path.isdir(self.session_temp_folder)):
                self.status_update.emit(f"Deleting temporary folder: {self.session_temp_folder}")
                self.logger.log("Deleting temporary files.")
                try:
                    shutil.rmtree(self.session_temp_folder)
                    self.status_update.emit("Temporary files deleted.")
                    self.logger.log("Temporary files deleted successfully.")
                except Exception as exc:
                    self.logger.log(f"ERROR: Could not delete temp folder: {exc}")
                    self.error_signal.emit(f"Could not delete temp folder: {exc}")
            if self.error_occurred:
                self.status_update.emit("Processing failed due to an error. Temporary files have been preserved for inspection.")
                self.logger.log("Run failed due to an error.")
            elif self._is_running:
                self.status_update.emit("Processing complete!")
                self.logger.log("Run completed successfully.")
            else:
                self.status_update.emit("Processing stopped.")
                self.logger.log("Run was stopped by the user.")
            self.logger.log_total_time()
            self.finished_signal.emit()

    def stop_processing(self):
        self._is_running = False
