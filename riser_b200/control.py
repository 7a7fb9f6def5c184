"""Batched drop-in for riser/control.py's SequencerControl.

Same constructor, same ``start`` / ``target`` / ``finish`` methods, same CSV rows,
counters, cache reset and ReadUntil calls (riser/control.py:4-153); only the serial
per-read body of the loop (control.py:31-93) is replaced by ONE call to
``BatchedClassifier.classify_batch`` per ``client.get_read_batch()``.
"""
import time

from .pipeline import BatchedClassifier, DECISION_NAMES, SKIPPED, ACCEPT, REJECT, NO_DECISION


class SequencerControl():
    def __init__(self, client, models, processor, logger, out_file, warm_up_batches=()):
        """Same arguments as riser/control.py:5.  ``warm_up_batches`` (additive, optional): batch sizes whose launch
        plans and CUDA graphs are built at the start of ``target`` instead of during the first polls that see them
        (e.g. ``(512,)`` for a MinION)."""
        self.client = client
        self.models = models
        self.proc = processor
        self.logger = logger
        self.out_filename = out_file
        self.classifier = BatchedClassifier(models, processor)
        self.batch_latencies = []      # seconds from "batch in hand" to "decisions on host"
        self.batch_sizes = []
        self.warm_up_batches = tuple(warm_up_batches)

    def target(self, mode, duration_h, threshold, unblock_duration=0.1):
        self.client.send_warning(
            'The sequencing run is being controlled by RISER, reads that are '
            'not in the target class will be ejected from the pore.')

        if self.warm_up_batches:
            self.classifier.warm_up(self.warm_up_batches, threshold, mode)

        with open(f'{self.out_filename}.csv', 'a') as out_file:
            self._write_header(out_file)
            run_start = time.monotonic()
            progress_time = run_start + 60
            duration_s = self._hours_to_seconds(duration_h)
            n_assessed = 0
            n_rejected = 0
            n_accepted = 0
            polyA_cache = {}
            while self.client.is_running() and time.monotonic() < run_start + duration_s:
                # Get batch of reads to process
                batch_start = time.monotonic()
                reads_to_reject = []
                reads_to_accept = []
                reads_unclassified = []
                batch = list(self.client.get_read_batch())
                signals = [self.client.get_raw_signal(read) for _, read in batch]
                t0 = time.monotonic()
                res = self.classifier.classify_batch(signals, [read.id for _, read in batch],
                                                     polyA_cache, threshold, mode)
                self.batch_latencies.append(time.monotonic() - t0)
                self.batch_sizes.append(len(batch))

                for i, (channel, read) in enumerate(batch):
                    code = int(res.decisions[i])
                    if code == SKIPPED:          # the `continue` branches, control.py:50,56
                        continue
                    n_assessed += 1
                    if code == ACCEPT:
                        reads_to_accept.append((channel, self._get_read_id(read)))
                    elif code == REJECT:
                        reads_to_reject.append((channel, self._get_read_id(read)))
                    elif code == NO_DECISION:
                        reads_unclassified.append((channel, self._get_read_id(read)))
                    self._write(out_file, batch_start, channel, read.id, int(res.sig_len[i]),
                                self.models, res.p_on[i], threshold, mode, DECISION_NAMES[code])

                # Send reject requests
                self.client.reject_reads(reads_to_reject, unblock_duration)
                n_rejected += len(reads_to_reject)

                # Don't need to reassess the reads that were rejected, accepted
                # or couldn't be classified after the maximum input length
                done = reads_to_reject + reads_to_accept + reads_unclassified
                self.client.finish_processing_reads(done)
                n_accepted += len(reads_to_accept)

                # Log progress each minute
                if batch_start > progress_time:
                    self.logger.info(f"In the last minute {n_assessed} signals "
                                     f"were assessed, {n_accepted} were "
                                     f"accepted and {n_rejected} were rejected")
                    n_assessed = 0
                    n_rejected = 0
                    n_accepted = 0
                    progress_time = batch_start + 60
            else:
                self.client.send_warning('RISER has stopped running.')
                if not self.client.is_running():
                    self.logger.info('Client has stopped.')
                if time.monotonic() > run_start + duration_s:
                    self.logger.info(f'RISER has timed out after {duration_h} '
                                     'hours as requested.')

    def start(self):
        self.client.start_streaming_reads()
        self.logger.info('Live read stream started.')

    def finish(self):
        self.client.reset()
        self.logger.info('Client reset and live read stream ended.')

    def _hours_to_seconds(self, hours):
        return hours * 60 * 60

    def _get_read_id(self, read):
        # Support for minknow-api <= v5.*
        if hasattr(read, "number"):
            return read.number
        # Support for minknow-api >= v6.*
        else:
            return read.id

    def _write_header(self, csv_file):
        csv_file.write('batch_start,read_id,channel,sig_length,models,prob_targets,threshold,mode,decision\n')

    def _write(self, csv_file, batch_start, channel, read, sig_length,
               models, p_on_targets, threshold, mode, decision):
        csv_file.write(f'{batch_start:.0f},{read},{channel},{sig_length},'
                       f'{";".join([m.target for m in models])},'
                       f'{";".join([str(float(p)) for p in p_on_targets])},'
                       f'{threshold},{mode},{decision}\n')
