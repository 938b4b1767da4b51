"""CPU evaluation glue around the GPU path: the reference's path-log
format and its beat-error score, so that field logs and CSV ground truth can score GPU paths.

Reference: log reader ``tests.py:12-27`` (5 header lines, then ``"<live> <ref>"``), log writer
``livenote_live.py:138-154`` / ``wtw_live.py:169-174`` (``\\r\\n`` line ends), scorer
``tests.py:29-137`` (``test_simple.get_error / get_beat / get_time / get_secs_off``;
frame -> seconds uses ``2048 / 22050.``).  This is host-side bookkeeping, not part of the hot path.
"""
import csv


def write_path_log(filename, ref_name, fft_len, hop_size, params, path):
    """livenote_live.py:138-154: five header lines, then one "live ref" pair per line (CRLF)."""
    keys = list(params.items())
    assert len(keys) == 2, "the reference logs exactly two algorithm parameters"
    with open(filename, "w", newline="") as f:
        f.write("%s\r\n" % ref_name)
        f.write("fft_len: %d\r\n" % fft_len)
        f.write("hop_size: %d\r\n" % hop_size)
        for k, v in keys:
            f.write("%s: %d\r\n" % (k, v))
        for l, r in path:
            f.write("%d %d\r\n" % (l, r))


def read_path_log(filename, header_lines=5):
    """tests.py:20-27 (data_from_file): skip the header, parse "live ref" pairs."""
    path = []
    with open(filename) as f:
        for line in f.readlines()[header_lines:]:
            tok = line.strip().split("\t")[0].split(" ")
            if len(tok) >= 2 and tok[0] != "":
                path.append((int(tok[0]), int(tok[1])))
    return path


def read_beats_csv(filename):
    """Songs/<piece>/<rec>.csv: rows time_seconds,beat_index[,label] (tests.py:46-56)."""
    times, beats = [], []
    with open(filename) as f:
        for row in csv.reader(f):
            if not row:
                continue
            times.append(float(row[0]))
            beats.append(int(row[1]))
    return times, beats


class BeatScorer(object):
    """test_simple (tests.py:29-137): % of path points whose live/ref beat positions differ by
    more than 1/3/5/10 beats or seconds.  Quirks kept: points whose interpolated beat is 0 or
    outside the annotated range are skipped (`if l_beat and r_beat`); seconds are looked up in the
    LIVE recording's beat times for both sides (get_time)."""

    def __init__(self, ref_csv, live_csv):
        self.ref_gt_times, self.ref_gt_beats = read_beats_csv(ref_csv)
        self.live_gt_times, self.live_gt_beats = read_beats_csv(live_csv)

    @staticmethod
    def get_beat(sample, gt_times, gt_beats):
        time = sample * (2048 / 22050.)
        for i in range(len(gt_times)):
            if i == 0:
                if time <= gt_times[i]:
                    frac = float(gt_times[i] - time) / (gt_times[i] - 0) if gt_times[i] != 0 else 0
                    return gt_beats[i] - frac
            elif gt_times[i - 1] <= time <= gt_times[i]:
                frac = float(gt_times[i] - time) / (gt_times[i] - gt_times[i - 1])
                return gt_beats[i] - frac
        return None

    def get_time(self, beat):
        t = self.live_gt_times[int(beat)]
        if int(beat) + 1 < len(self.live_gt_times):
            t += (beat % 1) * (self.live_gt_times[int(beat) + 1] - self.live_gt_times[int(beat)])
        return t

    def score(self, path):
        thresholds = (1, 3, 5, 10)
        off_beats = dict.fromkeys(thresholds, 0)
        off_secs = dict.fromkeys(thresholds, 0)
        count = 0
        for l, r in path:
            l_beat = self.get_beat(l, self.live_gt_times, self.live_gt_beats)
            r_beat = self.get_beat(r, self.ref_gt_times, self.ref_gt_beats)
            if l_beat and r_beat:
                diff = abs(l_beat - r_beat)
                secs = abs(self.get_time(r_beat) - self.get_time(l_beat))
                for th in thresholds:
                    off_beats[th] += diff > th
                    off_secs[th] += secs > th
                count += 1
        if count == 0:
            return {"count": 0}
        out = {"count": count}
        for th in thresholds:
            out["pct_off_%d_beats" % th] = float(off_beats[th]) / count * 100
            out["pct_off_%d_secs" % th] = float(off_secs[th]) / count * 100
        return out

    def get_error(self, path):
        """The number test_simple.get_error() returns: % of points off by more than 3 seconds."""
        return self.score(path).get("pct_off_3_secs")
