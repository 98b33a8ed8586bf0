"""
uvtools_wrapper -- host process boundary to UVToolsCmd (reference uvtools_wrapper.py:6-100).  Not part of
the device path; kept so the "uvtools" input mode of ProcessingPipelineThread keeps working: same
command lines, same accepted exit codes (0 and 1), same .uvtop layout and output naming.
"""
import os
import re
import subprocess

_TRAILING_INT = re.compile(r"(\d+)\.\w+$")


def _layer_number(name):
    m = _TRAILING_INT.search(name)
    return int(m.group(1)) if m else float("inf")


def _run_uvtools(command, what):
    flags = subprocess.CREATE_NO_WINDOW if os.name == "nt" else 0
    try:
        proc = subprocess.run(command, capture_output=True, text=True, creationflags=flags)
    except Exception as exc:
        raise RuntimeError(f"An unexpected error occurred during UVTools {what}: {exc}")
    if proc.returncode not in (0, 1):   # 1 is a benign exit code of UVToolsCmd
        raise RuntimeError(f"UVTools exited with an unexpected error (code {proc.returncode}):\n\n{proc.stderr}")


def extract_layers(uvtools_path: str, input_file: str, temp_folder: str) -> str:
    input_folder = os.path.join(temp_folder, "Input")
    os.makedirs(input_folder, exist_ok=True)
    _run_uvtools([uvtools_path, "extract", input_file, input_folder, "--content", "Layers"], "extraction")
    return input_folder


def generate_uvtop_file(processed_images_folder: str, temp_folder: str, run_timestamp: str) -> str:
    files = sorted((os.path.join(processed_images_folder, f) for f in os.listdir(processed_images_folder)
                    if f.lower().endswith(".png")), key=_layer_number)
    if not files:
        raise RuntimeError("No processed image files found to generate .uvtop file.")
    lines = ['<?xml version="1.0" encoding="utf-8" standalone="no"?>',
             '<OperationLayerImport xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
             'xmlns:xsd="http://www.w3.org/2001/XMLSchema">',
             "  <LayerRangeSelection>None</LayerRangeSelection>", "  <ImportType>Replace</ImportType>", "  <Files>"]
    for path in files:
        lines += ["    <GenericFileRepresentation>", f"      <FilePath>{path}</FilePath>",
                  "    </GenericFileRepresentation>"]
    lines += ["  </Files>", "</OperationLayerImport>", ""]
    uvtop = os.path.join(temp_folder, f"repack_operations_{run_timestamp}.uvtop")
    with open(uvtop, "w", encoding="utf-8") as fh:
        fh.write("\n".join(lines))
    return uvtop


def repack_layers(uvtools_path: str, input_file: str, uvtop_filepath: str, output_location: str, temp_folder: str,
                  output_prefix: str, run_timestamp: str) -> str:
    out_dir = os.path.dirname(input_file) if output_location == "input_folder" else temp_folder
    final_path = os.path.join(out_dir, f"{output_prefix}{run_timestamp}_{os.path.basename(input_file)}")
    _run_uvtools([uvtools_path, "run", input_file, uvtop_filepath, "--output", final_path], "repacking")
    return final_path
