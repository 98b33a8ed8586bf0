"""logger -- the run log of the pipeline thread: one file per run under logs/, named by the start time, one
"[time] message" line per event (reference logger.py:5-36: same constructor, attributes and method names, so
processing_pipeline and ui_components use it unchanged)."""
import datetime as _dt
import json
import pathlib


def _now():
    return _dt.datetime.now()


class Logger:
    FILE_STAMP = "%Y%m%d-%H%M%S"
    LINE_STAMP = "%Y-%m-%d %H:%M:%S.%f"

    def __init__(self, log_dir="logs"):
        folder = pathlib.Path(log_dir)
        folder.mkdir(parents=True, exist_ok=True)
        self.start_time = _now()
        self.run_timestamp = self.start_time.strftime(self.FILE_STAMP)
        self.log_filepath = str(folder / (self.run_timestamp + ".log"))

    def _write(self, message: str):
        with open(self.log_filepath, "a", encoding="utf-8") as out:
            print(message, file=out)

    def log(self, message: str):
        millis = _now().strftime(self.LINE_STAMP)[:-3]          # microseconds cut to milliseconds
        self._write("[" + millis + "] " + str(message))

    def log_config(self, config_obj):
        self.log("Starting run with the following configuration:")
        try:
            text = json.dumps(config_obj.to_dict(), indent=2)
        except Exception as problem:
            self.log(f"Could not serialize config object: {problem}")
        else:
            self._write(text)
        self.log("-" * 20)

    def log_total_time(self):
        self.log(f"Total execution time: {_now() - self.start_time}")
