"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the month glue of the reference's
`oisatgmi` class, the calls run/job.py:61-84 makes after `read_data`:

  recal_amf     /root/reference/oisatgmi/driver.py:36-39
  cal_pwv       /root/reference/oisatgmi/driver.py:41-43
  conv_ak       /root/reference/oisatgmi/driver.py:45-51
  average       /root/reference/oisatgmi/driver.py:53-63   (O3: model column -> DU)
  bias_correct  /root/reference/oisatgmi/driver.py:65-106
  oi            /root/reference/oisatgmi/driver.py:108-114 (GOSAT: OI on aux2 / aux1)

Parity status: PINNED -- tests/test_oracle_vs_reference.py runs the same month
through the unmodified class and compares bit for bit (the knee inside OI is the
one exception, see oracle/kneedle.py).  Only tests/, smoke() and bench.py's CPU
legs import this.
"""
from __future__ import annotations

from oracle import averaging as _avg, oi as _oi, vertical as _vert


class oisatgmi(object):

    def attach(self, reader_obj, gasname):
        self.reader_obj = reader_obj
        self.gasname = gasname

    def recal_amf(self):
        self.reader_obj.sat_data = _vert.amf_recal(self.reader_obj.ctm_data,
                                                   self.reader_obj.sat_data)

    def cal_pwv(self):
        from oracle import ssmis as _ssmis
        self.reader_obj.sat_data = _ssmis.pwv_calculator(self.reader_obj.ctm_data,
                                                         self.reader_obj.sat_data)

    def conv_ak(self, sensor):
        if sensor == "MOPITT":
            self.reader_obj.sat_data = _vert.ak_conv_mopitt(self.reader_obj.ctm_data,
                                                            self.reader_obj.sat_data)
        if sensor == "GOSAT":
            self.reader_obj.sat_data = _vert.ak_conv_gosat(self.reader_obj.ctm_data,
                                                           self.reader_obj.sat_data)

    def average(self, startdate, enddate, gasname=None):
        (self.sat_averaged_vcd, self.sat_averaged_error, self.ctm_averaged_vcd, self.aux1,
         self.aux2, self.avg_time) = _avg.averaging(startdate, enddate, self.reader_obj)
        if gasname == "O3":
            self.ctm_averaged_vcd = self.ctm_averaged_vcd / (2.69e16 * 1e-15)

    def bias_correct(self, sat_type, gasname):
        self.sat_averaged_vcd = _oi.bias_correct(sat_type, gasname, self.sat_averaged_vcd)

    def oi(self, sensor, error_ctm=50.0):
        res = _oi.oi_from_means(sensor, self.ctm_averaged_vcd, self.sat_averaged_vcd,
                                self.sat_averaged_error, self.aux1, self.aux2, error_ctm)
        (self.ctm_averaged_vcd_corrected, self.ak_OI, self.increment_OI, self.error_OI) = res[:4]
        self.knee_index = res[4]
