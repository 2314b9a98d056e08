"""Build profiles/ncu_traffic.json from `ncu --set full ... --page raw --csv` exports of profiles/prof_small.py runs.

    python profiles/make_traffic.py spectral:512:<main_frames>:<raw.csv> all:1024:<main_frames>:<raw.csv>

For every kernel group of bench.py (the names afx_batch_kernel_times reports) the DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) of the group's launches are summed and divided by the main frames of the captured compute."""
import csv, json, os, sys

GROUPS = {
    "spectrum": ["k_spectrum", "k_flux"],
    "rhythm": ["k_rhythm_front", "k_rhythm_polar", "k_rhythm_pipe", "k_rhythm_whiten", "k_rhythm_odf", "k_rhythm_power", "k_rhythm_median", "k_rhythm_back"],
    "pitch": ["k_pitch", "k_pitch_hop"],
    "bands": ["k_bands_a_big", "k_bands_a_small", "k_bands_b", "k_bands_select", "k_bands_lane"],
    "autocorr": ["k_autocorr"],
    "peaks": ["k_whiten_main", "k_peaks_count", "k_peaks_file", "k_peaks_pipe"],
    "stats": ["k_stats"],
    "condition": ["k_downmix", "k_trim", "k_eff"],
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    kn, rd, wr = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    out = {}
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").split("<")[0].strip()
        b = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
        out[name] = out.get(name, 0.0) + b
    return out


def main():
    res = {"how": "ncu --set full --clock-control none, one compute inside an NVTX range (profiles/gpu_prof.sh, "
                  "profiles/prof_small.py); dram__bytes_read.sum + dram__bytes_write.sum per kernel, divided by the main "
                  "frames of the launch; rebuilt by profiles/make_traffic.py"}
    for spec in sys.argv[1:]:
        feat, hop, frames, path = spec.split(":")
        per_kernel = load(path)
        groups = {}
        for g, ks in GROUPS.items():
            have = [k for k in ks if k in per_kernel]
            if have:
                groups[g] = {"dram_bytes_per_main_frame": sum(per_kernel[k] for k in have) / float(frames), "kernels": have}
        res[feat] = {"hop": int(hop), "main_frames": int(frames), "groups": groups, "source": os.path.relpath(path)}
    json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps({k: {g: round(v["dram_bytes_per_main_frame"]) for g, v in res[k]["groups"].items()} for k in res if k != "how"}))


if __name__ == "__main__":
    main()
