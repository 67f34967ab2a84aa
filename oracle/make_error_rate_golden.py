"""TEST INFRASTRUCTURE — extracts the known answers for EditStatistics.word_error_rate and its accumulation
from the reference's published result files (/root/reference/interspeech_results/*.json) into a small fixture.

    python -m oracle.make_error_rate_golden
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main() -> None:
    entries, totals = [], []
    for path in sorted(glob.glob("/root/reference/interspeech_results/*.json")):
        results = json.load(open(path))["results"]
        languages = {k: v for k, v in results.items() if k != "total"}
        classifiers = list(next(iter(languages.values()))["error_statistics"])
        for classifier in classifiers:
            rows = []
            for language, record in results.items():
                stats = record["error_statistics"].get(classifier)
                rate = record.get("error_rates", {}).get(classifier)
                if stats is None:
                    continue
                row = [stats["insertions"], stats["deletions"], stats["substitutions"], stats["correct"]]
                if rate is not None:
                    entries.append(row + [rate])
                if language != "total":
                    rows.append(row)
            if "total" in results and classifier in results["total"]["error_statistics"]:
                total = results["total"]["error_statistics"][classifier]
                totals.append([rows, [total["insertions"], total["deletions"], total["substitutions"], total["correct"]]])
    out = os.path.join(ROOT, "tests", "golden", "interspeech_error_rates.json")
    json.dump({"source": "interspeech_results/*.json (package_version 0.7.7)", "entries": entries, "totals": totals}, open(out, "w"))
    print(f"{len(entries)} error rates, {len(totals)} totals -> {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
