"""logger -- timestamped run log in logs/ (reference logger.py:5-36), same file naming and line format."""
import datetime
import json
import os


class Logger:
    def __init__(self, log_dir="logs"):
        os.makedirs(log_dir, exist_ok=True)
        self.start_time = datetime.datetime.now()
        self.run_timestamp = self.start_time.strftime("%Y%m%d-%H%M%S")
        self.log_filepath = os.path.join(log_dir, f"{self.run_timestamp}.log")

    def _write(self, message: str):
        with open(self.log_filepath, "a", encoding="utf-8") as fh:
            fh.write(message + "\n")

    def log(self, message: str):
        stamp = datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")[:-3]
        self._write(f"[{stamp}] {message}")

    def log_config(self, config_obj):
        self.log("Starting run with the following configuration:")
        try:
            self._write(json.dumps(config_obj.to_dict(), indent=2))
        except Exception as exc:
            self.log(f"Could not serialize config object: {exc}")
        self.log("-" * 20)

    def log_total_time(self):
        self.log(f"Total execution time: {datetime.datetime.now() - self.start_time}")
